// b2048_agent_inst.cu -- one agent size per translation unit: nvcc -DB2048_N=<2..6> (see 2048_b200/build.py).
#include "b2048_agent.cuh"

#ifndef B2048_N
#error "compile with -DB2048_N=2..6"
#endif
#define B2048_CAT2(a, b) a##b
#define B2048_CAT(a, b) B2048_CAT2(a, b)

extern const b2048_agent_ops B2048_CAT(b2048_agent_ops_, B2048_N) = {
    features_impl<B2048_N>,  evaluate_impl<B2048_N>,   td_update_impl<B2048_N>,
    greedy_play_impl<B2048_N>, td_phase_a_impl<B2048_N>, td_run_persistent<B2048_N>,
    look_forward_impl<B2048_N>, expectimax_play_impl<B2048_N>,
};
