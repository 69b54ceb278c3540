// agogo_b200 — dual.Train on device (K7/K8) interface.
#pragma once
#include "nn.cuh"

struct TrainWS { void* impl = nullptr; };
void train_ws_alloc(TrainWS& ws, const NetLayout& L);
void train_ws_free(TrainWS& ws);
float* train_ws_grads(TrainWS& ws);  // device, Model() layout
float* train_ws_cost(TrainWS& ws);   // device scalar
void train_ws_inputs(TrainWS& ws, float** X, float** Pi, float** V);  // device staging of one batch
void train_step_grads(TrainWS& ws, const NetLayout& L, const float* params, cudaStream_t st, unsigned long long* launches);
void train_sgd(TrainWS& ws, const NetLayout& L, float* params, float lr, float gscale, cudaStream_t st, unsigned long long* launches);
// K8: one kernel = reduce-scatter over peer memory + SGD on the owned slice + all-gather into every peer
void train_allreduce_sgd_p2p(float* const* peer_grads, float* const* peer_params, int* const* peer_flags, int* my_flags,
                             int rank, int world, size_t n, float lr, int epoch, unsigned int* done_counter, int num_sms,
                             int* err, cudaStream_t st, unsigned long long* launches);
