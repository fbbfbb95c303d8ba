// generated: thread-per-chain kernels for dim=10
#include "ptg_inst.cuh"
PTG_INSTANTIATE(10)
