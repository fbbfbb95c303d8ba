// generated: thread-per-chain kernels for dim=14
#include "ptg_inst.cuh"
PTG_INSTANTIATE(14)
