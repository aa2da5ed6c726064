// k_msm_g1_bls377.cu -- MSM / fixed-base kernels over G1 of BLS377
#include "msm.cuh"
namespace g16 {
G16_MSM_TEMPLATES(template, Fp<BLS377_FqP>, Fp<BLS377_FrP>)
}  // namespace g16
