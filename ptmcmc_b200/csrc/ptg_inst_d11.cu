// generated: thread-per-chain kernels for dim=11
#include "ptg_inst.cuh"
PTG_INSTANTIATE(11)
