// stand-in: its presence (with <Eigen/Sparse>) switches deviceflatassembler.hh to the Ikarus branch
#pragma once
#include "dirichletbcenforcement.hh"
