// k_msm_g1_bn254.cu -- MSM / fixed-base kernels over G1 of BN254
#include "msm.cuh"
namespace g16 {
G16_MSM_TEMPLATES(template, Fp<BN254_FqP>, Fp<BN254_FrP>)
}  // namespace g16
