// generated: thread-per-chain kernels for dim=8
#include "ptg_inst.cuh"
PTG_INSTANTIATE(8)
