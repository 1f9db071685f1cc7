// Standalone GPU check of the tcgen05 tap-GEMM kernels (pddm_conv2d_fwd / pddm_conv2d_wgrad) against a plain
// CPU evaluation of the contract written in include/pddm.h.  No torch: starts in milliseconds on the GPU box.
//   build: see tests/cuda/Makefile-less recipe in __graft_entry__.build()    run: ./test_conv
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/pddm.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static bool g_verify = true;
static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}
static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

struct Case {
  const char* name;
  int x_NB, B, H, W, Cin, ldx, xoff, Cout, ntaps;
  int db[9], dh[9], dw[9];
  int out_H, out_W, sh, sw, oh, ow;
  int use_bias, use_bcast, res_dtype /* -1 none */, y_dtype;
};

static void taps3x3(Case& c) {
  c.ntaps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      c.db[r * 3 + s] = 0;
      c.dh[r * 3 + s] = r - 1;
      c.dw[r * 3 + s] = s - 1;
    }
}

template <class F>
static float time_in_graph(F launch, cudaEvent_t e0, cudaEvent_t e1) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < 10; ++i) launch(s);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(s);
  return ms / 30;
}

static int run_case(const Case& c) {
  const size_t nx = (size_t)c.x_NB * c.H * c.W * c.ldx;
  const size_t nw = (size_t)c.Cout * c.ntaps * c.Cin;
  const size_t ny = (size_t)c.B * c.out_H * c.out_W * c.Cout;
  std::vector<float> x(nx), w(nw), bias(c.Cout), bcast((size_t)c.B * c.Cout), res(ny), yref(ny, 0.f);
  for (auto& v : x) v = bf(frand());
  for (auto& v : w) v = bf(frand() * 0.1f);
  for (auto& v : bias) v = frand();
  for (auto& v : bcast) v = frand();
  for (auto& v : res) v = c.res_dtype == PDDM_BF16 ? bf(frand()) : frand();
  std::vector<char> written(ny / c.Cout, 0);
  for (int b = 0; b < c.B && g_verify; ++b)
    for (int h = 0; h < c.H; ++h)
      for (int ww = 0; ww < c.W; ++ww) {
        const size_t opix = ((size_t)b * c.out_H + (h * c.sh + c.oh)) * c.out_W + (ww * c.sw + c.ow);
        written[opix] = 1;
        for (int n = 0; n < c.Cout; ++n) {
          double acc = 0;
          for (int t = 0; t < c.ntaps; ++t) {
            const int bb = b + c.db[t], hh = h + c.dh[t], w2 = ww + c.dw[t];
            if (hh < 0 || hh >= c.H || w2 < 0 || w2 >= c.W || bb < 0 || bb >= c.x_NB) continue;
            const float* xp = &x[(((size_t)bb * c.H + hh) * c.W + w2) * c.ldx + c.xoff];
            const float* wp = &w[((size_t)n * c.ntaps + t) * c.Cin];
            for (int k = 0; k < c.Cin; ++k) acc += (double)xp[k] * wp[k];
          }
          if (c.use_bias) acc += bias[n];
          if (c.use_bcast) acc += bcast[(size_t)b * c.Cout + n];
          if (c.res_dtype >= 0) acc += res[opix * c.Cout + n];
          yref[opix * c.Cout + n] = (float)acc;
        }
      }
  std::vector<__nv_bfloat16> xh(nx), wh(nw), resh(ny);
  for (size_t i = 0; i < nx; ++i) xh[i] = __float2bfloat16(x[i]);
  for (size_t i = 0; i < nw; ++i) wh[i] = __float2bfloat16(w[i]);
  for (size_t i = 0; i < ny; ++i) resh[i] = __float2bfloat16(res[i]);
  void *dx, *dw_, *dres, *dy;
  float *dbias, *dbcast;
  const size_t ybytes = ny * (c.y_dtype == PDDM_BF16 ? 2 : 4);
  CK(cudaMalloc(&dx, nx * 2));
  CK(cudaMalloc(&dw_, nw * 2));
  CK(cudaMalloc(&dres, ny * 4));
  CK(cudaMalloc(&dy, ybytes));
  CK(cudaMalloc(&dbias, c.Cout * 4));
  CK(cudaMalloc(&dbcast, (size_t)c.B * c.Cout * 4));
  CK(cudaMemcpy(dx, xh.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw_, wh.data(), nw * 2, cudaMemcpyHostToDevice));
  if (c.res_dtype == PDDM_BF16) CK(cudaMemcpy(dres, resh.data(), ny * 2, cudaMemcpyHostToDevice));
  else CK(cudaMemcpy(dres, res.data(), ny * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbcast, bcast.data(), (size_t)c.B * c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dy, 0, ybytes));

  pddm_conv_params p;
  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)dx + c.xoff;
  p.w = dw_;
  p.bias = c.use_bias ? dbias : nullptr;
  p.bcast = c.use_bcast ? dbcast : nullptr;
  p.ld_bcast = c.Cout;
  p.residual = c.res_dtype >= 0 ? dres : nullptr;
  p.res_dtype = c.res_dtype >= 0 ? c.res_dtype : 0;
  p.y = dy;
  p.y_dtype = c.y_dtype;
  p.x_NB = c.x_NB; p.B = c.B; p.H = c.H; p.W = c.W; p.Cin = c.Cin; p.ldx = c.ldx; p.Cout = c.Cout;
  p.ntaps = c.ntaps;
  for (int t = 0; t < c.ntaps; ++t) { p.tap_db[t] = c.db[t]; p.tap_dh[t] = c.dh[t]; p.tap_dw[t] = c.dw[t]; p.tap_w[t] = t; }
  p.w_ntaps = c.ntaps;
  p.out_H = c.out_H; p.out_W = c.out_W; p.out_sh = c.sh; p.out_sw = c.sw; p.out_oh = c.oh; p.out_ow = c.ow;
  int rc = pddm_conv2d_fwd(&p, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc != 0 || e != cudaSuccess) {
    printf("[FAIL] %-28s rc=%d (%s) cuda=%s\n", c.name, rc, pddm_strerror(rc), cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> y(ny);
  if (c.y_dtype == PDDM_BF16) {
    std::vector<__nv_bfloat16> yh(ny);
    CK(cudaMemcpy(yh.data(), dy, ny * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ny; ++i) y[i] = __bfloat162float(yh[i]);
  } else {
    CK(cudaMemcpy(y.data(), dy, ny * 4, cudaMemcpyDeviceToHost));
  }
  double maxerr = 0, maxref = 0;
  size_t bad = 0;
  for (size_t i = 0; i < ny && g_verify; ++i) {
    if (!written[i / c.Cout]) {
      if (y[i] != 0.f) ++bad;  // pixels outside the output mapping must stay untouched
      continue;
    }
    maxref = fmax(maxref, fabs(yref[i]));
    const double err = fabs(y[i] - yref[i]);
    maxerr = fmax(maxerr, err);
    const double tol = (c.y_dtype == PDDM_BF16 ? 1e-2 : 2e-4) * fmax(1.0, fabs(yref[i]));
    if (!(err <= tol)) ++bad;
  }
  // forward timing: 10 launches captured in a CUDA graph (the product path replays graphs, and direct launches of
  // the small shapes would time the host-side tensor-map encode instead of the kernel)
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = time_in_graph([&](cudaStream_t s) { pddm_conv2d_fwd(&p, s); }, e0, e1);
  const double fl = 2.0 * c.B * c.H * c.W * c.Cout * c.ntaps * c.Cin;
  printf("[%s] fwd   %-28s maxerr=%.3e maxref=%.2f bad=%zu  %.1f us  %.1f TF/s\n", bad ? "FAIL" : " ok ", c.name, maxerr,
         maxref, bad, ms * 1e3, fl / ms * 1e-9);

  // ---------------- wgrad: dy := random bf16 in tile space [B,H,W,Cout] (only when the output map is identity)
  int wbad = 0;
  if (c.sh == 1 && c.sw == 1 && c.oh == 0 && c.ow == 0 && c.out_H == c.H && c.out_W == c.W) {
    std::vector<float> g(ny);
    std::vector<__nv_bfloat16> gh(ny);
    for (size_t i = 0; i < ny; ++i) { g[i] = bf(frand()); gh[i] = __float2bfloat16(g[i]); }
    std::vector<double> dwref(nw, 0.0);
    for (int b = 0; b < c.B && g_verify; ++b)
      for (int h = 0; h < c.H; ++h)
        for (int ww = 0; ww < c.W; ++ww) {
          const float* gp = &g[(((size_t)b * c.H + h) * c.W + ww) * c.Cout];
          for (int t = 0; t < c.ntaps; ++t) {
            const int bb = b + c.db[t], hh = h + c.dh[t], w2 = ww + c.dw[t];
            if (hh < 0 || hh >= c.H || w2 < 0 || w2 >= c.W || bb < 0 || bb >= c.x_NB) continue;
            const float* xp = &x[(((size_t)bb * c.H + hh) * c.W + w2) * c.ldx + c.xoff];
            for (int n = 0; n < c.Cout; ++n) {
              double* d = &dwref[((size_t)n * c.ntaps + t) * c.Cin];
              const double gv = gp[n];
              for (int k = 0; k < c.Cin; ++k) d[k] += gv * xp[k];
            }
          }
        }
    void* dg; float* ddw; void* ws;
    CK(cudaMalloc(&dg, ny * 2));
    CK(cudaMalloc(&ddw, nw * 4));
    CK(cudaMemcpy(dg, gh.data(), ny * 2, cudaMemcpyHostToDevice));
    pddm_wgrad_params q;
    memset(&q, 0, sizeof(q));
    q.x = p.x; q.dy = dg; q.dw = ddw;
    q.x_NB = c.x_NB; q.B = c.B; q.H = c.H; q.W = c.W; q.Cin = c.Cin; q.ldx = c.ldx; q.Cout = c.Cout; q.lddy = c.Cout;
    q.ntaps = c.ntaps;
    for (int t = 0; t < c.ntaps; ++t) { q.tap_db[t] = c.db[t]; q.tap_dh[t] = c.dh[t]; q.tap_dw[t] = c.dw[t]; }
    for (int layout = 0; layout < 2; ++layout) {
      q.dw_layout = layout;
      const size_t wsb = pddm_conv2d_wgrad_workspace(&q);
      CK(cudaMalloc(&ws, wsb > 0 ? wsb : 16));
      CK(cudaMemset(ddw, 0xFF, nw * 4));
      rc = pddm_conv2d_wgrad(&q, ws, wsb, nullptr);
      e = cudaDeviceSynchronize();
      if (rc != 0 || e != cudaSuccess) {
        printf("[FAIL] wgrad %-28s rc=%d (%s) cuda=%s\n", c.name, rc, pddm_strerror(rc), cudaGetErrorString(e));
        return 1;
      }
      std::vector<float> dwv(nw);
      CK(cudaMemcpy(dwv.data(), ddw, nw * 4, cudaMemcpyDeviceToHost));
      double werr = 0, wmax = 0;
      size_t nb = 0;
      for (int n = 0; n < c.Cout && g_verify; ++n)
        for (int t = 0; t < c.ntaps; ++t)
          for (int k = 0; k < c.Cin; ++k) {
            const double r = dwref[((size_t)n * c.ntaps + t) * c.Cin + k];
            const float v = layout == 0 ? dwv[((size_t)n * c.ntaps + t) * c.Cin + k]
                                        : dwv[((size_t)n * c.Cin + k) * c.ntaps + t];
            wmax = fmax(wmax, fabs(r));
            werr = fmax(werr, fabs(v - r));
            if (!(fabs(v - r) <= 2e-3 * fmax(1.0, fabs(r)) + 1e-3)) ++nb;
          }
      ms = time_in_graph([&](cudaStream_t s) { pddm_conv2d_wgrad(&q, ws, wsb, s); }, e0, e1);
      printf("[%s] wgrad %-28s layout=%d maxerr=%.3e maxref=%.2f bad=%zu ws=%.1fMB  %.1f us  %.1f TF/s\n",
             nb ? "FAIL" : " ok ", c.name, layout, werr, wmax, nb, wsb / 1e6, ms * 1e3, fl / ms * 1e-9);
      wbad += nb != 0;
      cudaFree(ws);
    }
    cudaFree(dg); cudaFree(ddw);
  }
  cudaFree(dx); cudaFree(dw_); cudaFree(dres); cudaFree(dy); cudaFree(dbias); cudaFree(dbcast);
  return (bad != 0) + wbad;
}

int main(int argc, char** argv) {
  int rc = pddm_check_device();
  printf("pddm version %d, device check: %s, SMs %d\n", pddm_version(), pddm_strerror(rc), pddm_sm_count());
  if (rc) return 3;
  const bool big = argc > 1 && !strcmp(argv[1], "big");
  g_verify = !big;
  std::vector<Case> cases;
  auto add3 = [&](const char* name, int B, int H, int W, int Cin, int Cout, int bias, int bc, int res, int yd) {
    Case c; memset(&c, 0, sizeof(c));
    c.name = name; c.x_NB = B; c.B = B; c.H = H; c.W = W; c.Cin = Cin; c.ldx = Cin; c.Cout = Cout;
    taps3x3(c);
    c.out_H = H; c.out_W = W; c.sh = c.sw = 1;
    c.use_bias = bias; c.use_bcast = bc; c.res_dtype = res; c.y_dtype = yd;
    cases.push_back(c);
  };
  auto add1 = [&](const char* name, int B, int H, int W, int Cin, int ldx, int xoff, int Cout, int yd) {
    Case c; memset(&c, 0, sizeof(c));
    c.name = name; c.x_NB = B; c.B = B; c.H = H; c.W = W; c.Cin = Cin; c.ldx = ldx; c.xoff = xoff; c.Cout = Cout;
    c.ntaps = 1; c.out_H = H; c.out_W = W; c.sh = c.sw = 1;
    c.use_bias = 1; c.res_dtype = -1; c.y_dtype = yd;
    cases.push_back(c);
  };
  if (!big) {
    add3("3x3 32x32 128->128 b2", 2, 32, 32, 128, 128, 1, 1, PDDM_BF16, PDDM_BF16);
    add3("3x3 16x16 256->256 b3", 3, 16, 16, 256, 256, 1, 0, PDDM_F32, PDDM_F32);
    add3("3x3 8x8 64->384 b5", 5, 8, 8, 64, 384, 0, 1, -1, PDDM_BF16);
    add3("3x3 4x4 512->256 b20", 20, 4, 4, 512, 256, 1, 1, PDDM_BF16, PDDM_BF16);
    add3("3x3 28x28 32->64 b2", 2, 28, 28, 32, 64, 1, 1, -1, PDDM_BF16);
    add3("3x3 16x16 128->96 b3 res", 3, 16, 16, 128, 96, 1, 1, PDDM_BF16, PDDM_BF16);  // swap path when forced
    add3("3x3 8x8 64->128 b9", 9, 8, 8, 64, 128, 1, 0, -1, PDDM_BF16);                   // tile spans 4 samples
    add3("3x3 14x14 96->64 b3", 3, 14, 14, 96, 64, 1, 0, PDDM_BF16, PDDM_F32);
    add3("3x3 7x7 128->64 b5", 5, 7, 7, 128, 64, 1, 0, -1, PDDM_BF16);
    add3("3x3 64x64 128->128 b1", 1, 64, 64, 128, 128, 1, 0, -1, PDDM_BF16);
    add3("3x3 16x16 32->32 b1", 1, 16, 16, 32, 32, 0, 0, -1, PDDM_F32);
    add1("1x1 16x16 384->256 b2", 2, 16, 16, 384, 384, 0, 256, PDDM_BF16);
    add1("1x1 slice ld512 off256", 2, 8, 8, 256, 512, 256, 128, PDDM_F32);
    add1("linear 200x512->1000", 1, 1, 200, 512, 512, 0, 1000, PDDM_F32);
    add1("qkv 1x1 16x16 256->768 b2", 2, 16, 16, 256, 256, 0, 768, PDDM_BF16);
    // many tiles per CTA, odd tile counts, several channel tiles
    add3("3x3 16x16 64->768 b25", 25, 16, 16, 64, 768, 1, 1, PDDM_BF16, PDDM_BF16);
    add1("1x1 8x8 128->2048 b37", 37, 8, 8, 128, 128, 0, 2048, PDDM_BF16);
    add3("3x3 8x8 96->640 b60", 60, 8, 8, 96, 640, 1, 0, -1, PDDM_BF16);
    {  // stride-2 as taps over a phase-split tensor [4B, H/2, W/2, C] + strided output mapping
      Case c; memset(&c, 0, sizeof(c));
      c.name = "phase taps + out stride"; c.B = 2; c.x_NB = 8; c.H = 8; c.W = 8; c.Cin = 64; c.ldx = 64; c.Cout = 64;
      c.ntaps = 9;
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          const int a = (r == 1) ? 0 : 1, bq = (s == 1) ? 0 : 1;
          c.db[r * 3 + s] = (a * 2 + bq) * c.B;
          c.dh[r * 3 + s] = (r == 0) ? -1 : 0;
          c.dw[r * 3 + s] = (s == 0) ? -1 : 0;
        }
      c.out_H = 16; c.out_W = 16; c.sh = 2; c.sw = 2; c.oh = 1; c.ow = 0;
      c.use_bias = 1; c.res_dtype = -1; c.y_dtype = PDDM_F32;
      cases.push_back(c);
    }
  } else {
    add3("3x3 32x32 128->128 b128", 128, 32, 32, 128, 128, 1, 1, -1, PDDM_BF16);
    add3("3x3 16x16 256->256 b128", 128, 16, 16, 256, 256, 1, 1, -1, PDDM_BF16);
    add3("3x3 16x16 512->256 b128", 128, 16, 16, 512, 256, 1, 1, -1, PDDM_BF16);
    add3("3x3 8x8 256->256 b128", 128, 8, 8, 256, 256, 1, 1, -1, PDDM_BF16);
    add3("3x3 4x4 256->256 b128", 128, 4, 4, 256, 256, 1, 1, -1, PDDM_BF16);
    add1("1x1 16x16 256->768 b128", 128, 16, 16, 256, 256, 0, 768, PDDM_BF16);
    add1("1x1 16x16 256->256 b128", 128, 16, 16, 256, 256, 0, 256, PDDM_BF16);
    add1("1x1 32x32 256->128 b128", 128, 32, 32, 256, 256, 0, 128, PDDM_BF16);
    add3("3x3 16x16 256->256 b128 +res", 128, 16, 16, 256, 256, 1, 1, PDDM_BF16, PDDM_BF16);
  }
  // optional: `big <case index>` runs a single case (for ncu captures)
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  int fails = 0;
  for (size_t i = 0; i < cases.size(); ++i)
    if (only < 0 || static_cast<int>(i) == only) fails += run_case(cases[i]);
  printf("%s: %d failing checks\n", fails ? "FAILED" : "ALL OK", fails);
  return fails ? 1 : 0;
}
