// generated: thread-per-chain kernels for dim=16
#include "ptg_inst.cuh"
PTG_INSTANTIATE(16)
