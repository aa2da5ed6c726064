// engine_bn254.cu -- instantiates the proving engine (NTT + MSM kernels, host orchestration) for BN254.
#include "engine.cuh"
namespace g16 {
IEngine* make_engine_bn254(int device, int* rc) { return make_engine<BN254_Params>(device, rc); }
}  // namespace g16
