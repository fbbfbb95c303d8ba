// generated: thread-per-chain kernels for dim=6
#include "ptg_inst.cuh"
PTG_INSTANTIATE(6)
