// generated: thread-per-chain kernels for dim=1
#include "ptg_inst.cuh"
PTG_INSTANTIATE(1)
