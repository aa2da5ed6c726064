// engine_bls381.cu -- instantiates the proving engine (NTT + MSM kernels, host orchestration) for BLS381.
#include "engine.cuh"
namespace g16 {
IEngine* make_engine_bls381(int device, int* rc) { return make_engine<BLS381_Params>(device, rc); }
}  // namespace g16
