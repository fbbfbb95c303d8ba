// generated: thread-per-chain kernels for dim=12
#include "ptg_inst.cuh"
PTG_INSTANTIATE(12)
