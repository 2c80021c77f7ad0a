// stand-in for ikarus/assembler/dirichletbcenforcement.hh
#pragma once
namespace Ikarus {
enum class DBCOption { Raw, Reduced, Full };
}
