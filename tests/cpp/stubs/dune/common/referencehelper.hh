// stand-in for <dune/common/referencehelper.hh>
#pragma once
#include "exceptions.hh"
