// generated: thread-per-chain kernels for dim=15
#include "ptg_inst.cuh"
PTG_INSTANTIATE(15)
