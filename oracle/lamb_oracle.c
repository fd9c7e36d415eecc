/* CPU oracle: plain-C restatement of the reference fused LAMB step.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never linked into the product.
 *
 * Follows (paths relative to /root/reference/PyTorch/LanguageModeling/BERT/):
 *   lamb_amp_opt/fused_lamb/fused_lamb.py:130-260        host sequence (found_inf, scale, norms, step++)
 *   lamb_amp_opt/csrc/multi_tensor_l2norm_kernel.cu:99-108,126-150   L2 norms + non-finite -> noop
 *   lamb_amp_opt/csrc/multi_tensor_lamb.cu:63-79         noop early-out, bias correction, clip factor
 *   lamb_amp_opt/csrc/multi_tensor_lamb.cu:121-157       stage 1 (moments, update)
 *   lamb_amp_opt/csrc/multi_tensor_lamb.cu:274-282,319-323  stage 2 (trust ratio, apply)
 *   lamb_amp_opt/csrc/multi_tensor_lamb.cu:406-408       beta3 = grad_averaging ? 1-beta1 : 1
 *
 * Element math is fp32 exactly as the kernel's MATH_T=float (multi_tensor_lamb.cu:40);
 * the norms are accumulated in double (the ground truth the fp32 block-reductions of
 * any GPU implementation approximate) and rounded to float once.
 *
 * Parity pinning: the reference ships no test vectors for LAMB; this file is
 * pinned on the GPU box against the reference's own CUDA kernels rebuilt from
 * /root/reference into oracle/_ref (tests/test_lamb_gpu.py::test_oracle_vs_reference_kernel)
 * and by the closed-form known-answer cases in tests/test_lamb_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* sum of squares in double; returns 1 if any element is non-finite */
int lamb_oracle_sumsq_f32(const float* x, int64_t n, double* out)
{
    double s = 0.0; int bad = 0;
#pragma omp parallel for reduction(+:s) reduction(|:bad) if (n > 65536)
    for (int64_t i = 0; i < n; ++i) { float f = x[i]; if (!isfinite(f)) bad = 1; s += (double)f * (double)f; }
    *out = s; return bad;
}

/* Stage 1 for one tensor (multi_tensor_lamb.cu:121-157).  g is already fp32 (16-bit grads are
 * widened by the caller, exactly as `r_g = l_g * inv_scale` widens them).  Writes update (fp32),
 * m, v.  Returns sum(update^2) in double. */
double lamb_oracle_stage1_f32(const float* g, const float* p, float* m, float* v, float* update,
                              int64_t n, float beta1, float beta2, float beta3,
                              float beta1_correction, float beta2_correction, float epsilon,
                              int mode, float decay, float clipped_global_grad_norm, float inv_scale)
{
    double usq = 0.0;
#pragma omp parallel for reduction(+:usq) if (n > 65536)
    for (int64_t i = 0; i < n; ++i) {
        float r_g = g[i] * inv_scale;
        float r_p = (decay == 0.0f) ? 0.0f : p[i];
        float r_m = m[i], r_v = v[i], upd;
        if (mode == 0) {             /* MOMENT_MODE_0: L2 regularisation */
            float sg = r_g / clipped_global_grad_norm;
            sg = sg + decay * r_p;
            r_m = r_m * beta1 + beta3 * sg;
            r_v = r_v * beta2 + (1.0f - beta2) * sg * sg;
            float mu = r_m / beta1_correction, vu = r_v / beta2_correction;
            upd = mu / (sqrtf(vu) + epsilon);
        } else {                     /* MOMENT_MODE_1: decoupled weight decay (adam_w_mode) */
            float sg = r_g / clipped_global_grad_norm;
            r_m = r_m * beta1 + beta3 * sg;
            r_v = r_v * beta2 + (1.0f - beta2) * sg * sg;
            float mu = r_m / beta1_correction, vu = r_v / beta2_correction;
            upd = (mu / (sqrtf(vu) + epsilon)) + (decay * r_p);
        }
        m[i] = r_m; v[i] = r_v; update[i] = upd;
        usq += (double)upd * (double)upd;
    }
    return usq;
}

/* Stage 2 for one tensor (multi_tensor_lamb.cu:274-323). */
void lamb_oracle_stage2_f32(float* p, const float* update, int64_t n, float lr,
                            float param_norm, float update_norm, float decay, int use_nvlamb)
{
    float ratio = lr;
    if (use_nvlamb || decay != 0.0f)
        ratio = (update_norm != 0.0f && param_norm != 0.0f) ? lr * (param_norm / update_norm) : lr;
#pragma omp parallel for if (n > 65536)
    for (int64_t i = 0; i < n; ++i) p[i] = p[i] - ratio * update[i];
}
