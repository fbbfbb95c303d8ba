// generated: thread-per-chain kernels for dim=4
#include "ptg_inst.cuh"
PTG_INSTANTIATE(4)
