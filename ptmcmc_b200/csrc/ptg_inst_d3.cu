// generated: thread-per-chain kernels for dim=3
#include "ptg_inst.cuh"
PTG_INSTANTIATE(3)
