// Fused elementwise diffusion math (fp32, NCHW): q_sample, per-sample squared error (+grad), the reverse
// p_sample step, variational-bound terms, timestep embedding.  HBM-bound: 128-bit vector accesses where the
// per-sample size allows, one pass over each operand.  Arithmetic follows the reference's op order with
// explicitly rounded (non-contracted) multiplies/adds so the elementwise results are bit-identical to the
// fp32 torch expressions they replace.
#include <cuda_bf16.h>

#include "host_common.h"

namespace pddm {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in thread 0 (blockDim.x <= 1024)
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// ------------------------------------------------------------------------------------------ q_sample
// src/engine.py:251-261: mean = x*a[t-1]; std = s[t-1]; x_t = mean + noise*std
template <int VEC>
__global__ void q_sample_kernel(pddm_q_sample_params p) {
  pdl_entry();
  const int per = p.chw / VEC;
  const long long total = static_cast<long long>(p.B) * per;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per);
    const int t = p.t ? static_cast<int>(p.t[b]) : p.t_const;
    const float a = __ldg(p.alphas_hat_sqrt + t - 1), s = __ldg(p.one_min_alphas_hat_sqrt + t - 1);
    if (VEC == 4) {
      const float4 x = reinterpret_cast<const float4*>(p.x0)[i];
      const float4 n = reinterpret_cast<const float4*>(p.noise)[i];
      float4 o;
      o.x = __fadd_rn(__fmul_rn(x.x, a), __fmul_rn(n.x, s));
      o.y = __fadd_rn(__fmul_rn(x.y, a), __fmul_rn(n.y, s));
      o.z = __fadd_rn(__fmul_rn(x.z, a), __fmul_rn(n.z, s));
      o.w = __fadd_rn(__fmul_rn(x.w, a), __fmul_rn(n.w, s));
      reinterpret_cast<float4*>(p.x_t)[i] = o;
    } else {
      p.x_t[i] = __fadd_rn(__fmul_rn(p.x0[i], a), __fmul_rn(p.noise[i], s));
    }
  }
}

// ------------------------------------------------------------------------------------------ squared error
// one block per sample: L_b = mean_{c<C,hw} (noise - pred)^2 ; optional gradient wrt pred
__global__ void sq_err_kernel(pddm_sq_err_params p) {
  pdl_entry();
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const int n = p.C * p.hw;
  const float* pr = p.pred + static_cast<size_t>(b) * p.c_total * p.hw;  // first C channels are contiguous
  const float* nz = p.noise + static_cast<size_t>(b) * n;
  float* gp = p.grad_pred ? p.grad_pred + static_cast<size_t>(b) * p.c_total * p.hw : nullptr;
  const float gs = gp ? p.gscale[b] * 2.0f / static_cast<float>(n) : 0.f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = nz[i] - pr[i];
    acc += d * d;
    if (gp) gp[i] = -gs * d;
  }
  if (gp && p.grad_v_unit) {
    const float vs = p.gscale[b] * p.v_scale;
    const float* gv = p.grad_v_unit + static_cast<size_t>(b) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) gp[n + i] = vs * gv[i];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0 && p.per_sample) p.per_sample[b] = acc / static_cast<float>(n);
}

// ------------------------------------------------------------------------------------------ p_sample
// src/engine.py:348-397.  x0hat = A*x_t - Bc*eps (clamped), mean = c1*x0hat + c2*x_t   (clip)
//                         mean = (x_t - eps*dc) / sqrt(alpha)                            (no clip)
//                         x_{t-1} = mean - sigma*z   (z skipped at t == 1)
template <int VEC>
__global__ void p_sample_kernel(pddm_p_sample_params p) {
  pdl_entry();
  const int t = p.t_step_dev ? *p.t_step_dev : p.t_step;
  const int chw = p.C * p.hw;
  const int per = chw / VEC;
  const long long total = static_cast<long long>(p.B) * per;
  const float rA = __ldg(p.tab.sqrt_recip_alphas_cumprod + t - 1);
  const float rB = __ldg(p.tab.sqrt_recipm1_alphas_cumprod + t - 1);
  const float c1 = __ldg(p.tab.posterior_mean_coef1 + t - 1);
  const float c2 = __ldg(p.tab.posterior_mean_coef2 + t - 1);
  const float dc = __ldg(p.tab.denoising_coef + t - 1);
  const float as = __ldg(p.tab.alphas_sqrt + t - 1);
  float sigma = 0.f, min_log = 0.f, max_log = 0.f;
  if (p.sigma_mode == 0) sigma = __fsqrt_rn(__ldg(p.tab.betas + t - 1));
  else if (p.sigma_mode == 1) sigma = __fsqrt_rn(__ldg(p.tab.posterior_variance + t - 1));
  else {
    min_log = __ldg(p.tab.posterior_log_variance_clipped + t - 1);
    max_log = __ldg(p.tab.log_betas + t - 1);
  }
  const bool add_noise = (p.z != nullptr) && (t > 1);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per);
    const int r = static_cast<int>(i - static_cast<long long>(b) * per) * VEC;  // offset inside the sample
    const size_t xo = static_cast<size_t>(b) * chw + r;
    const size_t mo = static_cast<size_t>(b) * p.c_out * p.hw + r;
    float xt[VEC], ep[VEC], zz[VEC], vv[VEC], out[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(xt) = *reinterpret_cast<const float4*>(p.x_t + xo);
      *reinterpret_cast<float4*>(ep) = *reinterpret_cast<const float4*>(p.model_out + mo);
      if (add_noise) *reinterpret_cast<float4*>(zz) = *reinterpret_cast<const float4*>(p.z + xo);
      if (p.sigma_mode == 2) *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(p.model_out + mo + chw);
    } else {
      xt[0] = p.x_t[xo];
      ep[0] = p.model_out[mo];
      if (add_noise) zz[0] = p.z[xo];
      if (p.sigma_mode == 2) vv[0] = p.model_out[mo + chw];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float mean;
      if (p.clip) {
        float x0h = __fsub_rn(__fmul_rn(rA, xt[j]), __fmul_rn(rB, ep[j]));
        x0h = fminf(fmaxf(x0h, -1.f), 1.f);
        mean = __fadd_rn(__fmul_rn(x0h, c1), __fmul_rn(xt[j], c2));
      } else {
        mean = __fdiv_rn(__fsub_rn(xt[j], __fmul_rn(ep[j], dc)), as);
      }
      if (add_noise) {
        float s = sigma;
        if (p.sigma_mode == 2) {
          const float frac = (vv[j] + 1.f) * 0.5f;
          s = expf(0.5f * (frac * max_log + (1.f - frac) * min_log));
        }
        mean = __fsub_rn(mean, __fmul_rn(s, zz[j]));
      }
      out[j] = mean;
    }
    if (VEC == 4) *reinterpret_cast<float4*>(p.x_prev + xo) = *reinterpret_cast<float4*>(out);
    else p.x_prev[xo] = out[0];
  }
}

__global__ void step_advance_kernel(int32_t* t_dev, float* t_vec, int B) {
  pdl_entry();
  // single block; every thread reads the old value before anyone writes the new one
  const int t_new = *t_dev - 1;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) t_vec[b] = static_cast<float>(t_new);
  if (threadIdx.x == 0) *t_dev = t_new;
}

// ------------------------------------------------------------------------------------------ VLB terms
__device__ __forceinline__ float approx_cdf(float x) {  // src/utils.py:80-85
  return 0.5f * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float approx_pdf(float x) {  // d/dx approx_cdf
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float th = tanhf(u);
  return 0.5f * (1.f - th * th) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * x * x);
}
// log-likelihood of a discretised Gaussian (src/utils.py:88-115) and its derivative wrt log_scale
__device__ __forceinline__ float disc_loglik(float x, float mean, float log_scale, float* dlogp_dls) {
  const float cx = x - mean;
  const float inv = expf(-log_scale);
  const float pin = inv * (cx + 1.0f / 255.0f), min_ = inv * (cx - 1.0f / 255.0f);
  const float cp = approx_cdf(pin), cm = approx_cdf(min_);
  float val, d = 0.f;
  if (x < -0.999f) {
    val = logf(fmaxf(cp, 1e-12f));
    if (dlogp_dls && cp > 1e-12f) d = approx_pdf(pin) * (-pin) / cp;
  } else if (x > 0.999f) {
    const float om = 1.0f - cm;
    val = logf(fmaxf(om, 1e-12f));
    if (dlogp_dls && om > 1e-12f) d = approx_pdf(min_) * min_ / om;
  } else {
    const float delta = cp - cm;
    val = logf(fmaxf(delta, 1e-12f));
    if (dlogp_dls && delta > 1e-12f) d = (approx_pdf(pin) * (-pin) - approx_pdf(min_) * (-min_)) / delta;
  }
  if (dlogp_dls) *dlogp_dls = d;
  return val;
}
__device__ __forceinline__ float normal_kl(float m1, float lv1, float m2, float lv2) {  // src/utils.py:50-77
  const float dm = m1 - m2;
  return 0.5f * (-1.0f + lv2 - lv1 + expf(lv1 - lv2) + dm * dm * expf(-lv2));
}

// one block per sample
__global__ void vlb_kernel(pddm_vlb_params p) {
  pdl_entry();
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const int chw = p.C * p.hw;
  const float inv_ln2 = 1.4426950408889634f;
  const float* x0 = p.x0 + static_cast<size_t>(b) * chw;
  float acc = 0.f;
  if (p.mode == 2) {  // L_T
    const float a = __ldg(p.tab.alphas_hat_sqrt + p.tab.T - 1);
    const float s = __ldg(p.tab.one_min_alphas_hat_sqrt + p.tab.T - 1);
    const float lv1 = 2.f * logf(s);
    for (int i = threadIdx.x; i < chw; i += blockDim.x) acc += normal_kl(x0[i] * a, lv1, 0.f, 0.f);
  } else {
    const int t = static_cast<int>(p.t[b]);
    const float* xt = p.x_t + static_cast<size_t>(b) * chw;
    const float* mo = p.model_out + static_cast<size_t>(b) * p.c_out * p.hw;
    const float c1 = __ldg(p.tab.posterior_mean_coef1 + t - 1), c2 = __ldg(p.tab.posterior_mean_coef2 + t - 1);
    if (p.mode == 0) {
      const float dc = __ldg(p.tab.denoising_coef + t - 1), as = __ldg(p.tab.alphas_sqrt + t - 1);
      const float var = p.sigma_mode == 0 ? __ldg(p.tab.betas + t - 1) : __ldg(p.tab.posterior_variance + t - 1);
      const float sigma = sqrtf(var);
      const float plogvar = 2.f * logf(sigma), logscale = logf(sigma);
      const float qlogvar = logf(__ldg(p.tab.posterior_variance + t - 1));
      for (int i = threadIdx.x; i < chw; i += blockDim.x) {
        const float pmean = (xt[i] - mo[i] * dc) / as;
        if (t == 1) acc -= disc_loglik(x0[i], pmean, logscale, nullptr);
        else acc += normal_kl(x0[i] * c1 + xt[i] * c2, qlogvar, pmean, plogvar);
      }
    } else {  // learned variance
      const float rA = __ldg(p.tab.sqrt_recip_alphas_cumprod + t - 1);
      const float rB = __ldg(p.tab.sqrt_recipm1_alphas_cumprod + t - 1);
      const float min_log = __ldg(p.tab.posterior_log_variance_clipped + t - 1);
      const float max_log = __ldg(p.tab.log_betas + t - 1);
      const float gfac = 0.5f * (max_log - min_log) * inv_ln2 / static_cast<float>(chw);
      float* gv = p.grad_v ? p.grad_v + static_cast<size_t>(b) * chw : nullptr;
      for (int i = threadIdx.x; i < chw; i += blockDim.x) {
        const float x0h = rA * xt[i] - rB * mo[i];
        const float pmean = x0h * c1 + xt[i] * c2;
        const float frac = (mo[chw + i] + 1.f) * 0.5f;
        const float logvar = frac * max_log + (1.f - frac) * min_log;
        float g;
        if (t == 1) {
          float dls;
          acc -= disc_loglik(x0[i], pmean, 0.5f * logvar, gv ? &dls : nullptr);
          g = gv ? -0.5f * dls : 0.f;
        } else {
          const float tm = x0[i] * c1 + xt[i] * c2;
          acc += normal_kl(tm, min_log, pmean, logvar);
          const float dm = tm - pmean;
          g = 0.5f * (1.f - expf(min_log - logvar) - dm * dm * expf(-logvar));
        }
        if (gv) gv[i] = g * gfac;
      }
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) p.out[b] = acc / static_cast<float>(chw) * inv_ln2;
}

// ------------------------------------------------------------------------------------------ timestep embedding
__global__ void temb_kernel(const void* t, int t_is_float, void* out, int out_dtype, int B, int dim, float neg_log_mp) {
  pdl_entry();
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim, j = idx - b * dim;
  const float tv = t_is_float ? static_cast<const float*>(t)[b] : static_cast<float>(static_cast<const int64_t*>(t)[b]);
  float v = 0.f;
  if (j < 2 * half) {
    const int i = j < half ? j : j - half;
    const float f = expf(__fdiv_rn(__fmul_rn(neg_log_mp, static_cast<float>(i)), static_cast<float>(half)));
    const float a = __fmul_rn(tv, f);
    v = j < half ? cosf(a) : sinf(a);
  }
  if (out_dtype == PDDM_BF16) static_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16(v);
  else static_cast<float*>(out)[idx] = v;
}

// ------------------------------------------------------------------------------------------ fused Adam (+EMA)
// torch.optim.Adam semantics (no amsgrad): m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ;
// p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps) ;  ema = d*ema + (1-d)*p
__global__ void adam_kernel(pddm_adam_params p, float bc1, float bc2_sqrt) {
  pdl_entry();
  const long long n4 = p.n / 4;
  if (p.step_dev) {
    const float st = static_cast<float>(*p.step_dev);
    bc1 = 1.f - powf(p.beta1, st);
    bc2_sqrt = sqrtf(1.f - powf(p.beta2, st));
  }
  const float step_size = (p.lr_dev ? *p.lr_dev : p.lr) / bc1;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 w = reinterpret_cast<float4*>(p.param)[i];
    float4 g = reinterpret_cast<const float4*>(p.grad)[i];
    float4 m = reinterpret_cast<float4*>(p.exp_avg)[i];
    float4 v = reinterpret_cast<float4*>(p.exp_avg_sq)[i];
    float* wp = &w.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gj = gp[j] * p.grad_scale;
      if (p.weight_decay != 0.f) gj += p.weight_decay * wp[j];
      mp[j] = p.beta1 * mp[j] + (1.f - p.beta1) * gj;
      vp[j] = p.beta2 * vp[j] + (1.f - p.beta2) * gj * gj;
      wp[j] -= step_size * mp[j] / (sqrtf(vp[j]) / bc2_sqrt + p.eps);
    }
    reinterpret_cast<float4*>(p.param)[i] = w;
    reinterpret_cast<float4*>(p.exp_avg)[i] = m;
    reinterpret_cast<float4*>(p.exp_avg_sq)[i] = v;
    if (p.ema) {
      float4 e = reinterpret_cast<float4*>(p.ema)[i];
      e.x = p.ema_decay * e.x + (1.f - p.ema_decay) * w.x;
      e.y = p.ema_decay * e.y + (1.f - p.ema_decay) * w.y;
      e.z = p.ema_decay * e.z + (1.f - p.ema_decay) * w.z;
      e.w = p.ema_decay * e.w + (1.f - p.ema_decay) * w.w;
      reinterpret_cast<float4*>(p.ema)[i] = e;
    }
  }
}

// multi-tensor variant: block -> (tensor, chunk) through a table; float4 path when all four arrays are 16 B aligned
__global__ void __launch_bounds__(256) adam_multi_kernel(const pddm_adam_tensor* __restrict__ descs,
                                                         const int2* __restrict__ blocks, pddm_adam_params p, float bc1,
                                                         float bc2_sqrt) {
  pdl_entry();
  const int2 blk = blocks[blockIdx.x];
  const pddm_adam_tensor d = descs[blk.x];
  if (p.step_dev) {
    const float st = static_cast<float>(*p.step_dev);
    bc1 = 1.f - powf(p.beta1, st);
    bc2_sqrt = sqrtf(1.f - powf(p.beta2, st));
  }
  const float step_size = (p.lr_dev ? *p.lr_dev : p.lr) / bc1;
  const float b1 = p.beta1, b2 = p.beta2, wd = p.weight_decay, gs = p.grad_scale, eps = p.eps, ed = p.ema_decay;
  const long long beg = blk.y;
  const long long end = min(static_cast<long long>(d.n), beg + PDDM_ADAM_CHUNK);
  auto upd = [&](float& w, float g, float& m, float& v) {
    g *= gs;
    if (wd != 0.f) g += wd * w;
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    w -= step_size * m / (sqrtf(v) / bc2_sqrt + eps);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(d.param) | reinterpret_cast<uintptr_t>(d.grad) |
                     reinterpret_cast<uintptr_t>(d.exp_avg) | reinterpret_cast<uintptr_t>(d.exp_avg_sq) |
                     reinterpret_cast<uintptr_t>(d.ema)) & 15) == 0;
  long long i = beg;
  if (vec) {
    const long long end4 = beg + ((end - beg) & ~3LL);
    for (i = beg + threadIdx.x * 4LL; i < end4; i += blockDim.x * 4LL) {
      float4 w = *reinterpret_cast<float4*>(d.param + i);
      const float4 g = *reinterpret_cast<const float4*>(d.grad + i);
      float4 m = *reinterpret_cast<float4*>(d.exp_avg + i);
      float4 v = *reinterpret_cast<float4*>(d.exp_avg_sq + i);
      upd(w.x, g.x, m.x, v.x);
      upd(w.y, g.y, m.y, v.y);
      upd(w.z, g.z, m.z, v.z);
      upd(w.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(d.param + i) = w;
      *reinterpret_cast<float4*>(d.exp_avg + i) = m;
      *reinterpret_cast<float4*>(d.exp_avg_sq + i) = v;
      if (d.ema) {
        float4 e = *reinterpret_cast<float4*>(d.ema + i);
        e.x = ed * e.x + (1.f - ed) * w.x;
        e.y = ed * e.y + (1.f - ed) * w.y;
        e.z = ed * e.z + (1.f - ed) * w.z;
        e.w = ed * e.w + (1.f - ed) * w.w;
        *reinterpret_cast<float4*>(d.ema + i) = e;
      }
    }
    i = end4;
  }
  for (long long k = i + threadIdx.x; k < end; k += blockDim.x) {  // unaligned tensors and the < 4 element tail
    float w = d.param[k], m = d.exp_avg[k], v = d.exp_avg_sq[k];
    upd(w, d.grad[k], m, v);
    d.param[k] = w;
    d.exp_avg[k] = m;
    d.exp_avg_sq[k] = v;
    if (d.ema) d.ema[k] = ed * d.ema[k] + (1.f - ed) * w;
  }
}

static int ew_grid(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(device_info().sm_count > 0 ? device_info().sm_count : 148) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace pddm

using namespace pddm;

extern "C" int pddm_q_sample(const pddm_q_sample_params* p, pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->x0 || !p->noise || !p->x_t || !p->alphas_hat_sqrt || !p->one_min_alphas_hat_sqrt || p->B <= 0 ||
      p->chw <= 0)
    return PDDM_ERR_BAD_ARG;
  const bool v4 = p->chw % 4 == 0 && aligned16(p->x0) && aligned16(p->noise) && aligned16(p->x_t);
  const long long work = static_cast<long long>(p->B) * p->chw / (v4 ? 4 : 1);
  if (v4) PdlLaunch(ew_grid(work, 256), 256, 0, s)(q_sample_kernel<4>, *p);
  else PdlLaunch(ew_grid(work, 256), 256, 0, s)(q_sample_kernel<1>, *p);
  return launch_status();
}

extern "C" int pddm_sq_err(const pddm_sq_err_params* p, pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->pred || !p->noise || p->B <= 0 || p->C <= 0 || p->hw <= 0 || p->c_total < p->C) return PDDM_ERR_BAD_ARG;
  if (p->grad_pred && !p->gscale) return PDDM_ERR_BAD_ARG;
  PdlLaunch(p->B, 256, 0, s)(sq_err_kernel, *p);
  return launch_status();
}

extern "C" int pddm_p_sample_step(const pddm_p_sample_params* p, pddm_stream_t s_) {
  cudaStream_t s = static_cast<cudaStream_t>(s_);
  if (!p || !p->x_t || !p->model_out || !p->x_prev || p->B <= 0 || p->C <= 0 || p->hw <= 0) return PDDM_ERR_BAD_ARG;
  if (p->sigma_mode < 0 || p->sigma_mode > 2) return PDDM_ERR_BAD_ARG;
  if (p->c_out != p->C && p->c_out != 2 * p->C) return PDDM_ERR_BAD_ARG;
  if (p->sigma_mode == 2 && p->c_out != 2 * p->C) return PDDM_ERR_BAD_ARG;
  if (!p->t_step_dev && (p->t_step < 1 || p->t_step > p->tab.T)) return PDDM_ERR_BAD_ARG;
  const int chw = p->C * p->hw;
  const bool v4 = chw % 4 == 0 && (p->c_out * p->hw) % 4 == 0 && aligned16(p->x_t) && aligned16(p->model_out) &&
                  aligned16(p->x_prev) && (!p->z || aligned16(p->z));
  const long long work = static_cast<long long>(p->B) * chw / (v4 ? 4 : 1);
  if (v4) PdlLaunch(ew_grid(work, 256), 256, 0, s)(p_sample_kernel<4>, *p);
  else PdlLaunch(ew_grid(work, 256), 256, 0, s)(p_sample_kernel<1>, *p);
  return launch_status();
}

extern "C" int pddm_step_advance(int32_t* t_dev, float* t_vec, int32_t B, pddm_stream_t s_) {
  if (!t_dev || !t_vec || B <= 0) return PDDM_ERR_BAD_ARG;
  PdlLaunch(1, 256, 0, static_cast<cudaStream_t>(s_))(step_advance_kernel, t_dev, t_vec, B);
  return launch_status();
}

extern "C" int pddm_vlb_terms(const pddm_vlb_params* p, pddm_stream_t s_) {
  if (!p || !p->x0 || !p->out || p->B <= 0 || p->C <= 0 || p->hw <= 0 || p->mode < 0 || p->mode > 2)
    return PDDM_ERR_BAD_ARG;
  if (p->mode != 2 && (!p->x_t || !p->model_out || !p->t)) return PDDM_ERR_BAD_ARG;
  if (p->mode == 1 && p->c_out != 2 * p->C) return PDDM_ERR_BAD_ARG;
  PdlLaunch(p->B, 256, 0, static_cast<cudaStream_t>(s_))(vlb_kernel, *p);
  return launch_status();
}

extern "C" int pddm_timestep_embedding(const void* t, int32_t t_is_float, void* out, int32_t out_dtype, int32_t B,
                                       int32_t dim, float max_period, pddm_stream_t s_) {
  if (!t || !out || B <= 0 || dim <= 0) return PDDM_ERR_BAD_ARG;
  const float neg_log_mp = -static_cast<float>(log(static_cast<double>(max_period)));
  const int n = B * dim;
  PdlLaunch((n + 255) / 256, 256, 0, static_cast<cudaStream_t>(s_))(temb_kernel, t, t_is_float, out, out_dtype, B, dim,
                                                                          neg_log_mp);
  return launch_status();
}

__global__ void counter_add_kernel(int32_t* c, int32_t d) {
  pdl_entry(); *c += d; }
extern "C" int pddm_counter_add(int32_t* counter, int32_t delta, pddm_stream_t s_) {
  if (!counter) return PDDM_ERR_BAD_ARG;
  PdlLaunch(1, 1, 0, static_cast<cudaStream_t>(s_))(counter_add_kernel, counter, delta);
  return launch_status();
}

extern "C" int pddm_adam_ema_step(const pddm_adam_params* p, pddm_stream_t s_) {
  if (!p || !p->param || !p->grad || !p->exp_avg || !p->exp_avg_sq || p->n <= 0 || p->n % 4 != 0 ||
      (!p->step_dev && p->step < 1))
    return PDDM_ERR_BAD_ARG;
  const float bc1 = 1.f - static_cast<float>(pow(static_cast<double>(p->beta1), p->step));
  const float bc2 = 1.f - static_cast<float>(pow(static_cast<double>(p->beta2), p->step));
  PdlLaunch(ew_grid(p->n / 4, 256), 256, 0, static_cast<cudaStream_t>(s_))(adam_kernel, *p, bc1, sqrtf(bc2));
  return launch_status();
}

extern "C" int pddm_adam_ema_multi(const void* descs, const void* blocks, int32_t nblocks, const pddm_adam_params* p,
                                   pddm_stream_t s_) {
  if (!descs || !blocks || !p || nblocks <= 0 || p->step < 0) return PDDM_ERR_BAD_ARG;
  if (!p->step_dev && p->step < 1) return PDDM_ERR_BAD_ARG;
  const float bc1 = 1.f - static_cast<float>(pow(static_cast<double>(p->beta1), p->step));
  const float bc2 = 1.f - static_cast<float>(pow(static_cast<double>(p->beta2), p->step));
  PdlLaunch(nblocks, 256, 0, static_cast<cudaStream_t>(s_))(adam_multi_kernel,
                                                             static_cast<const pddm_adam_tensor*>(descs),
                                                             static_cast<const int2*>(blocks), *p, bc1, sqrtf(bc2));
  return launch_status();
}
