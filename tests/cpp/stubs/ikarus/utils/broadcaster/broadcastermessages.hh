// stand-in for ikarus/utils/broadcaster/broadcastermessages.hh
#pragma once
namespace Ikarus {
enum class NonLinearSolverMessages { BEGIN, INIT, ITERATION_STARTED, ITERATION_ENDED, CORRECTION_UPDATED, FINISHED_SUCESSFULLY, END };
}
