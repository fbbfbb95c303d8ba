// generated: thread-per-chain kernels for dim=9
#include "ptg_inst.cuh"
PTG_INSTANTIATE(9)
