// generated: thread-per-chain kernels for dim=2
#include "ptg_inst.cuh"
PTG_INSTANTIATE(2)
