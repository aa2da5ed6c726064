// k_msm_g1_bls381.cu -- MSM / fixed-base kernels over G1 of BLS381
#include "msm.cuh"
namespace g16 {
G16_MSM_TEMPLATES(template, Fp<BLS381_FqP>, Fp<BLS381_FrP>)
}  // namespace g16
