// generated: thread-per-chain kernels for dim=7
#include "ptg_inst.cuh"
PTG_INSTANTIATE(7)
