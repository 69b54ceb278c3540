#include "tower_tc.cuh"
bool tc_tower_supported(const NetDims&) { return false; }
void tc_tower_alloc(TcTower&, const NetDims&, int, int) {}
void tc_tower_free(TcTower&) {}
void tc_tower_prepare(TcTower&, const NetLayout&, const Snapshot&, cudaStream_t, unsigned long long*) {}
void tc_tower_forward(TcTower&, const NetLayout&, const Snapshot&, Fp32Scratch&, const float*, const int*, int, float*, int,
                      float*, int*, cudaStream_t, unsigned long long*) {}
