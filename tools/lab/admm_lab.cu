// LAB (not part of the product): times k_admm<QpikCfg<7>> alone on QP records made by tools/lab/make_records.py.
//   admm_lab records.bin N_base [tile] : throughput run (records tiled to N_base*tile robots, schedule = input order)
//                                        + lone-warp latency run (1 block, tolerances 0 => runs to max_iter)
// Prints per-variant timing and a checksum of (iters, out) so that variants can be compared for equality.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dyros_robot_controller_b200/csrc/drc_kernels.cuh"
using namespace drc;
using namespace drc_kernels;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#ifndef LAB_MINB
#define LAB_MINB 3
#endif
typedef QpikCfg<7> Cfg;

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "tools/lab/records.bin";
  const int nb = argc > 2 ? atoi(argv[2]) : 4096, tile = argc > 3 ? atoi(argv[3]) : 16;
  const int B = nb * tile, ST = Cfg::STRIDE;
  std::vector<double> rec((size_t)nb * ST);
  FILE* f = fopen(path, "rb");
  if (!f || fread(rec.data(), sizeof(double), rec.size(), f) != rec.size()) { printf("cannot read %s\n", path); return 1; }
  fclose(f);
  double* d_qp; double* d_out; int *d_status, *d_iters;
  CK(cudaMalloc(&d_qp, (size_t)B * ST * sizeof(double)));
  for (int t = 0; t < tile; ++t) CK(cudaMemcpy(d_qp + (size_t)t * nb * ST, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, (size_t)B * 7 * sizeof(double)));
  CK(cudaMalloc(&d_status, B * sizeof(int))); CK(cudaMalloc(&d_iters, B * sizeof(int)));
  DrcParams prm;
  QpOptions o = qp_options(prm, (1u << 7) - 1u);
  SolveIO io;
  memset(&io, 0, sizeof io);
  io.B = B; io.qp = d_qp; io.out = d_out; io.sout = aos(7); io.status = d_status; io.iters = d_iters;
  constexpr size_t smem = sizeof(GroupShared<Cfg>) * kAdmmWarps * Cfg::NG;
  CK(cudaFuncSetAttribute(k_admm<Cfg, false, LAB_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_block = kAdmmWarps * Cfg::NG, blocks = (B + per_block - 1) / per_block;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(a));
    k_admm<Cfg, false, LAB_MINB><<<blocks, kAdmmWarps * 32, smem>>>(io, o);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<int> it(B), st(B); std::vector<double> out((size_t)B * 7);
  CK(cudaMemcpy(it.data(), d_iters, B * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st.data(), d_status, B * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(double), cudaMemcpyDeviceToHost));
  long long sumit = 0; int solved = 0, maxit = 0; double cs = 0;
  for (int i = 0; i < nb; ++i) { sumit += it[i]; solved += st[i] == 1; if (it[i] > maxit) maxit = it[i]; for (int k = 0; k < 7; ++k) cs += out[(size_t)i * 7 + k] * (1 + k + (i % 13)); }
  printf("throughput: B=%d  %.3f ms  (%.2f M robots/s)  mean iters %.3f max %d solved %d/%d checksum %.12e\n", B, best, B / best / 1e3,
         (double)sumit / nb, maxit, solved, nb, cs);
  // capped run: every robot stops at 200 iterations at the latest => no convergence tail, pure bulk throughput
  {
    QpOptions oc = o; oc.max_iter = 200;
    float bc = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(a));
      k_admm<Cfg, false, LAB_MINB><<<blocks, kAdmmWarps * 32, smem>>>(io, oc);
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < bc) bc = ms;
    }
    printf("bulk (max_iter 200): %.3f ms\n", bc);
  }
  // lone warp: one block, never converges
  {
    QpOptions ol = o; ol.eps_abs = 0; ol.eps_rel = 0; ol.eps_prim_inf = 0; ol.eps_dual_inf = 0; ol.adaptive_rho = 0;
    SolveIO il = io; il.B = Cfg::NG;
    float bl = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(a));
      k_admm<Cfg, false, LAB_MINB><<<1, kAdmmWarps * 32, smem>>>(il, ol);
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < bl) bl = ms;
    }
    int it0; CK(cudaMemcpy(&it0, d_iters, sizeof(int), cudaMemcpyDeviceToHost));
    printf("lone warp: %.3f ms for %d iterations = %.1f ns / iteration\n", bl, it0, bl * 1e6 / it0);
    {  // lone warp without termination checks: the hot loop alone
      QpOptions on = ol; on.check_termination = 0;
      float bn = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        k_admm<Cfg, false, LAB_MINB><<<1, kAdmmWarps * 32, smem>>>(il, on);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < bn) bn = ms;
      }
      printf("lone warp, no checks: %.3f ms = %.1f ns / iteration  => one check costs %.0f ns\n", bn, bn * 1e6 / 4000, (bl - bn) * 1e6 / 160);
      on.max_iter = 400;
      float bq = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        k_admm<Cfg, false, LAB_MINB><<<blocks, kAdmmWarps * 32, smem>>>(io, on);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < bq) bq = ms;
      }
      printf("full load, 400 iterations, no checks: %.3f ms\n", bq);
      on.max_iter = 25;
      float b0 = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        k_admm<Cfg, false, LAB_MINB><<<blocks, kAdmmWarps * 32, smem>>>(io, on);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < b0) b0 = ms;
      }
      printf("full load, 25 iterations, no checks (setup: load + Ruiz + factor): %.3f ms\n", b0);
      for (int sc : {0, 5}) {   // Ruiz passes: cost of the equilibration itself
        QpOptions os = on; os.scaling = sc;
        float bs = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(a));
          k_admm<Cfg, false, LAB_MINB><<<blocks, kAdmmWarps * 32, smem>>>(io, os);
          CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
          float ms; CK(cudaEventElapsedTime(&ms, a, b));
          if (rep > 0 && ms < bs) bs = ms;
        }
        printf("  same with %d Ruiz passes: %.3f ms\n", sc, bs);
      }
      {
        QpOptions os = on; os.max_iter = 1;
        float bs = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(a));
          k_admm<Cfg, false, LAB_MINB><<<blocks, kAdmmWarps * 32, smem>>>(io, os);
          CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
          float ms; CK(cudaEventElapsedTime(&ms, a, b));
          if (rep > 0 && ms < bs) bs = ms;
        }
        printf("  10 Ruiz passes, 1 iteration + final check: %.3f ms\n", bs);
      }
    }
    // all warps loaded, nobody converges, 400 iterations
    ol.max_iter = 400;
    float bf = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(a));
      k_admm<Cfg, false, LAB_MINB><<<blocks, kAdmmWarps * 32, smem>>>(io, ol);
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < bf) bf = ms;
    }
    printf("full load, 400 iterations each: %.3f ms = %.3f ns / robot-iteration\n", bf, bf * 1e6 / ((double)B * 400));
  }
  return 0;
}
