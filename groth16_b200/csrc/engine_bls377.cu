// engine_bls377.cu -- instantiates the proving engine (NTT + MSM kernels, host orchestration) for BLS377.
#include "engine.cuh"
namespace g16 {
IEngine* make_engine_bls377(int device, int* rc) { return make_engine<BLS377_Params>(device, rc); }
}  // namespace g16
