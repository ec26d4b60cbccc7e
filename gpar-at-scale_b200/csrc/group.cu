// group.cu — several devices of one box driven from ONE host process through the C ABI (SURVEY 8e, 8b
// `gpar_group_*`): the Julia package is a single process, so "one process per GPU + torch.distributed"
// (gpar-at-scale_b200/parallel.py, bench.py --gpus N) is not available to it.
//
// The shards are the ones the reference's drivers expose: per-output conditional GPs and hyper-parameter
// restarts (examples/GPAR_scaled_examples.jl:132-175 fits output i on the OBSERVED outputs < i, so the fits are
// independent; util.jl:128-134 draws the restart start points).  A group owns one gpar_ctx per device and
//   * evaluates one objective per member concurrently (one host thread per device: every entry point of the
//     single-device ABI blocks on its own stream),
//   * runs whole Nelder-Mead fits (dtc.jl:58-61; temporal_gp_inference.jl:82) of a task list, tasks handed out
//     dynamically, longest first,
//   * all-gathers the SCALARS (status, value, gradient / minimum, minimiser) with NCCL so that every device holds
//     the table, and broadcasts posterior means down the GPAR chain (GPAR_scaled_examples.jl:172) with ncclBroadcast
//     over NVLink.  No collective touches the data path of these task-sharded calls;
//   * evaluates ONE objective whose rows are sharded over the members (gpar_group_*_sharded): per-slice statistics, one or two
//     small all-gathers (filter carries) and one all-reduce of the M x M statistics per evaluation.
// NCCL is bound at run time (dlopen of libnccl.so.2: the library has no link-time dependency on it).
#include "common.cuh"
#include "optim_host.h"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <functional>
#include <cmath>
#include <limits>
#include <numeric>
#include <thread>

struct gpar_group {
  std::vector<gpar_ctx*> ctx;
  std::vector<int> dev;
  std::vector<ncclComm_t> comm;
  std::vector<DevBuf> send, recv;       // scalar tables
  void* lib = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool loopback = false;                // testing: members share ONE device, collectives are device copies / sums (no NCCL)
  std::string err;
};

namespace {

int group_fail(gpar_group* g, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (g) g->err = buf;
  return code;
}
#define GCU(call)                                                                                                   \
  do { cudaError_t e_ = (call);                                                                                     \
       if (e_ != cudaSuccess) return group_fail(g, GPAR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define GNC(call)                                                                                                   \
  do { ncclResult_t r_ = (call);                                                                                    \
       if (r_ != ncclSuccess) return group_fail(g, GPAR_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g->GetErrorString(r_), __FILE__, __LINE__); } while (0)

int load_nccl(gpar_group* g) {
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g->lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g->lib) break; }
  if (!g->lib) return group_fail(g, GPAR_ERR_CUDA, "gpar_group_create: libnccl.so.2 not found (%s)", dlerror());
#define BIND(field, sym)                                                                                            \
  g->field = reinterpret_cast<decltype(g->field)>(dlsym(g->lib, sym));                                               \
  if (!g->field) return group_fail(g, GPAR_ERR_CUDA, "gpar_group_create: symbol %s missing in libnccl", sym)
  BIND(CommInitAll, "ncclCommInitAll"); BIND(CommDestroy, "ncclCommDestroy"); BIND(AllGather, "ncclAllGather"); BIND(AllReduce, "ncclAllReduce");
  BIND(Broadcast, "ncclBroadcast"); BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd");
  BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  return GPAR_OK;
}

__global__ void add_into_kernel(double* __restrict__ acc, const double* __restrict__ x, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) acc[i] += x[i];
}
// The data-path collectives.  With NCCL: one grouped call over the members' streams.  Loopback groups (all members on one
// device — the 1-GPU test mode of the sharded entry points) do the same data movement with device copies.
int group_allgather(gpar_group* g, const std::vector<const double*>& send, const std::vector<double*>& recv, size_t count) {
  const int n = (int)g->ctx.size();
  if (!g->loopback) {
    GNC(g->GroupStart());
    for (int i = 0; i < n; i++) GNC(g->AllGather(send[i], recv[i], count, ncclDouble, g->comm[i], g->ctx[i]->stream));
    GNC(g->GroupEnd());
    return GPAR_OK;
  }
  for (int i = 0; i < n; i++) GCU(cudaStreamSynchronize(g->ctx[i]->stream));
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) GCU(cudaMemcpyAsync(recv[i] + (size_t)j * count, send[j], count * sizeof(double), cudaMemcpyDeviceToDevice, g->ctx[i]->stream));
  return GPAR_OK;
}
int group_allreduce_sum(gpar_group* g, const std::vector<double*>& buf, size_t count) {
  const int n = (int)g->ctx.size();
  if (!g->loopback) {
    GNC(g->GroupStart());
    for (int i = 0; i < n; i++) GNC(g->AllReduce(buf[i], buf[i], count, ncclDouble, ncclSum, g->comm[i], g->ctx[i]->stream));
    GNC(g->GroupEnd());
    return GPAR_OK;
  }
  for (int i = 0; i < n; i++) GCU(cudaStreamSynchronize(g->ctx[i]->stream));
  cudaStream_t s0 = g->ctx[0]->stream;
  for (int j = 1; j < n; j++) add_into_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s0>>>(buf[0], buf[j], count);
  GCU(cudaGetLastError());
  GCU(cudaStreamSynchronize(s0));
  for (int j = 1; j < n; j++) GCU(cudaMemcpyAsync(buf[j], buf[0], count * sizeof(double), cudaMemcpyDeviceToDevice, g->ctx[j]->stream));
  return GPAR_OK;
}

// rows: per member `width` doubles (host, member-major).  Every device ends up with the ndev x width table
// (g->recv[i]); `table` (host, nullable) receives member 0's copy.
int allgather_rows(gpar_group* g, const double* rows, int width, double* table) {
  const int n = (int)g->ctx.size();
  for (int i = 0; i < n; i++) {
    GCU(cudaSetDevice(g->dev[i]));
    GCU(g->send[i].reserve((size_t)width * sizeof(double)));
    GCU(g->recv[i].reserve((size_t)width * n * sizeof(double)));
    GCU(cudaMemcpyAsync(g->send[i].p, rows + (size_t)i * width, (size_t)width * sizeof(double), cudaMemcpyHostToDevice, g->ctx[i]->stream));
  }
  {
    std::vector<const double*> sp(n); std::vector<double*> rp(n);
    for (int i = 0; i < n; i++) { sp[i] = g->send[i].as<double>(); rp[i] = g->recv[i].as<double>(); }
    const int rc = group_allgather(g, sp, rp, (size_t)width);
    if (rc != GPAR_OK) return rc;
  }
  for (int i = 0; i < n; i++) {
    GCU(cudaSetDevice(g->dev[i]));
    if (i == 0 && table) GCU(cudaMemcpyAsync(table, g->recv[0].p, (size_t)width * n * sizeof(double), cudaMemcpyDeviceToHost, g->ctx[0]->stream));
    GCU(cudaStreamSynchronize(g->ctx[i]->stream));
  }
  return GPAR_OK;
}

// one host thread per member; fn(i) -> status of member i
template <class Fn>
void run_members(gpar_group* g, std::vector<int>& codes, Fn fn) {
  const int n = (int)g->ctx.size();
  codes.assign(n, GPAR_OK);
  if (n == 1) { codes[0] = fn(0); return; }
  std::vector<std::thread> th;
  for (int i = 0; i < n; i++) th.emplace_back([&, i]() { cudaSetDevice(g->dev[i]); codes[i] = fn(i); });
  for (auto& t : th) t.join();
}

// Member i evaluates thetas[:, i] on ITS OWN resident data (gpar_set_* on gpar_group_ctx(g, i)), all members at once.
// codes[i] = the member's status: a failed Cholesky (GPAR_ERR_NOT_POSDEF) does not fail the call — an optimiser treats
// it as +Inf — any other member error does.  width 8 rows (status, value, gradient) are all-gathered with NCCL.
int group_eval(gpar_group* g, int ntheta, const double* thetas, double* vals, double* grads, int32_t* codes,
                      const std::function<int(int, const double*, double*, double*)>& eval) {
  if (!g) return GPAR_ERR_INVALID;
  if (!thetas || !vals) return group_fail(g, GPAR_ERR_INVALID, "group evaluation: thetas and vals must not be NULL");
  const int n = (int)g->ctx.size();
  std::vector<double> rows((size_t)n * 8, 0.0), table((size_t)n * 8);
  std::vector<int> st;
  run_members(g, st, [&](int i) {
    double v = std::numeric_limits<double>::quiet_NaN(), gr[5] = {0, 0, 0, 0, 0};
    int rc = eval(i, thetas + (size_t)ntheta * i, &v, grads ? gr : nullptr);
    rows[(size_t)i * 8] = rc; rows[(size_t)i * 8 + 1] = v;
    for (int j = 0; j < ntheta; j++) rows[(size_t)i * 8 + 2 + j] = gr[j];
    return rc;
  });
  for (int i = 0; i < n; i++)
    if (st[i] != GPAR_OK && st[i] != GPAR_ERR_NOT_POSDEF) return group_fail(g, st[i], "member %d (device %d): %s", i, g->dev[i], gpar_last_error(g->ctx[i]));
  int rc = allgather_rows(g, rows.data(), 8, table.data());
  if (rc != GPAR_OK) return rc;
  for (int i = 0; i < n; i++) {
    if (codes) codes[i] = (int32_t)table[(size_t)i * 8];
    vals[i] = table[(size_t)i * 8 + 1];
    if (grads) for (int j = 0; j < ntheta; j++) grads[(size_t)ntheta * i + j] = table[(size_t)i * 8 + 2 + j];
  }
  return GPAR_OK;
}

}  // namespace

extern "C" {

int gpar_group_create(const int32_t* devices, int32_t ndev, gpar_group** out) {
  if (!out) return GPAR_ERR_INVALID;
  *out = nullptr;
  if (!devices || ndev < 1) return GPAR_ERR_INVALID;
  gpar_group* g = new gpar_group();
  auto bail = [&](int code) { for (gpar_ctx* c : g->ctx) gpar_ctx_destroy(c); if (g->lib) dlclose(g->lib); delete g; return code; };
  // GPAR_GROUP_LOOPBACK=1 (testing): several members on ONE device, collectives as device copies — the sharded entry
  // points then run their slice logic on a 1-GPU box; otherwise one member per device
  bool same = ndev > 1;
  for (int i = 1; i < ndev; i++) same = same && devices[i] == devices[0];
  if (same) { const char* e = getenv("GPAR_GROUP_LOOPBACK"); g->loopback = e && atoi(e) != 0; }
  for (int i = 0; i < ndev; i++) {
    for (int j = 0; j < i; j++) if (devices[j] == devices[i] && !g->loopback) return bail(GPAR_ERR_INVALID);
    gpar_ctx* c = nullptr;
    int rc = gpar_ctx_create(devices[i], &c);
    if (rc != GPAR_OK) return bail(rc);
    g->ctx.push_back(c); g->dev.push_back(devices[i]);
  }
  g->send.resize(ndev); g->recv.resize(ndev);
  if (g->loopback) { *out = g; return GPAR_OK; }
  if (load_nccl(g) != GPAR_OK) { fprintf(stderr, "%s\n", g->err.c_str()); return bail(GPAR_ERR_CUDA); }
  g->comm.resize(ndev);
  ncclResult_t r = g->CommInitAll(g->comm.data(), ndev, g->dev.data());
  if (r != ncclSuccess) { fprintf(stderr, "gpar_group_create: ncclCommInitAll failed: %s\n", g->GetErrorString(r)); g->comm.clear(); return bail(GPAR_ERR_CUDA); }
  *out = g;
  return GPAR_OK;
}

int gpar_group_destroy(gpar_group* g) {
  if (!g) return GPAR_OK;
  for (size_t i = 0; i < g->ctx.size(); i++) {
    cudaSetDevice(g->dev[i]);
    cudaStreamSynchronize(g->ctx[i]->stream);
    if (i < g->comm.size()) g->CommDestroy(g->comm[i]);
    g->send[i].release(); g->recv[i].release();
    gpar_ctx_destroy(g->ctx[i]);
  }
  if (g->lib) dlclose(g->lib);
  delete g;
  return GPAR_OK;
}

int32_t gpar_group_size(const gpar_group* g) { return g ? (int32_t)g->ctx.size() : 0; }
gpar_ctx* gpar_group_ctx(gpar_group* g, int32_t member) { return (g && member >= 0 && member < (int32_t)g->ctx.size()) ? g->ctx[member] : nullptr; }
const char* gpar_group_last_error(const gpar_group* g) { return g ? g->err.c_str() : "null group"; }

int gpar_group_dtc_logpdf(gpar_group* g, int kernel, const double* thetas, int vfe, double jitter, double* vals, double* grads, int32_t* codes) {
  return group_eval(g, 3, thetas, vals, grads, codes, [&](int i, const double* th, double* v, double* gr) {
    return gpar_dtc_logpdf(g->ctx[i], kernel, th, vfe, jitter, v, gr);
  });
}

int gpar_group_scaled_dtc(gpar_group* g, int k_time, int k_out, const double* thetas, double* vals, double* grads, int32_t* codes) {
  return group_eval(g, 5, thetas, vals, grads, codes, [&](int i, const double* th, double* v, double* gr) {
    return gr ? gpar_scaled_dtc_grad(g->ctx[i], k_time, k_out, th, v, gr) : gpar_scaled_dtc(g->ctx[i], k_time, k_out, th, v, nullptr);
  });
}

// ONE plain DTC / VFE objective whose data are sharded over the members by rows (SURVEY 8e "intra-output N-sharding"):
// every member evaluates the sufficient statistics G = Kuf Kfu, H, g, h, y'y of ITS slice (same pseudo-inputs on every
// member; slices may differ in length), ONE ncclAllReduce sums the M x M + ... buffer over NVLink, member 0 runs the tail.
// 8 (2 M^2 + 2 M + 1) bytes per evaluation (16.8 MB at M = 1024); the row-sharded entry points are the only ones with a data-path collective.
int gpar_group_dtc_logpdf_sharded(gpar_group* g, int kernel, const double theta[3], int vfe, double jitter, double* val, double* grad) {
  if (!g) return GPAR_ERR_INVALID;
  if (!theta || !val) return group_fail(g, GPAR_ERR_INVALID, "dtc_logpdf_sharded: theta and val must not be NULL");
  if (kernel < GPAR_EQ || kernel > GPAR_MATERN52) return group_fail(g, GPAR_ERR_INVALID, "dtc_logpdf_sharded: unknown kernel code %d", kernel);
  const int n = (int)g->ctx.size();
  int64_t Ntot = 0;
  for (int i = 0; i < n; i++) {
    if (g->ctx[i]->M != g->ctx[0]->M || g->ctx[i]->Dz != g->ctx[0]->Dz)
      return group_fail(g, GPAR_ERR_INVALID, "dtc_logpdf_sharded: member %d holds %lld pseudo-inputs of dimension %d, member 0 %lld of %d",
                        i, (long long)g->ctx[i]->M, g->ctx[i]->Dz, (long long)g->ctx[0]->M, g->ctx[0]->Dz);
    Ntot += g->ctx[i]->N;
  }
  const GpParams p = unpack_gp3(theta);
  const bool want_grad = grad != nullptr;
  gpar_ctx* c0 = g->ctx[0];
  GCU(cudaSetDevice(g->dev[0]));
  int rc = dtc_tail_prepare(c0, kernel, p, vfe, jitter, want_grad);      // cov(u), L_u, ... on member 0's side stream
  if (rc != GPAR_OK) return group_fail(g, rc, "member 0: %s", gpar_last_error(c0));
  bool ill = false;
  {
    // The collapsed statistic G = Kuf Kfu loses cond(cov(u)) eps ("Conditioning").  When cov(u) is poorly conditioned
    // every member factors it too and whitens ITS panels by L_u before the SYRK, exactly as the one-device entry point does,
    // so that the result does not depend on how the rows were sharded.
    TailBufs tb;
    rc = tail_layout(c0, want_grad, vfe, &tb);
    if (rc != GPAR_OK) return group_fail(g, rc, "member 0: %s", gpar_last_error(c0));
    double mm[2] = {1.0, 1.0};
    GCU(cudaMemcpyAsync(mm, tb.sc + 4, sizeof(mm), cudaMemcpyDeviceToHost, c0->stream2));
    GCU(cudaStreamSynchronize(c0->stream2));
    ill = gpar_needs_whitened_panel(mm);
  }
  std::vector<int> st;
  if (ill) {
    run_members(g, st, [&](int i) { return i == 0 ? GPAR_OK : dtc_tail_prepare(g->ctx[i], kernel, p, vfe, jitter, want_grad); });
    for (int i = 0; i < n; i++)
      if (st[i] != GPAR_OK) return group_fail(g, st[i], "member %d (device %d): %s", i, g->dev[i], gpar_last_error(g->ctx[i]));
  }
  std::vector<double*> stats(n, nullptr);
  std::vector<size_t> count(n, 0);
  run_members(g, st, [&](int i) { return dtc_slice_stats(g->ctx[i], kernel, p, want_grad, &stats[i], &count[i], ill ? vfe : -1); });
  for (int i = 0; i < n; i++)
    if (st[i] != GPAR_OK) return group_fail(g, st[i], "member %d (device %d): %s", i, g->dev[i], gpar_last_error(g->ctx[i]));
  rc = group_allreduce_sum(g, stats, count[0]);
  if (rc != GPAR_OK) return rc;
  GCU(cudaSetDevice(g->dev[0]));
  const int M = (int)c0->M, Mpad = (M + GPAR_TILE - 1) / GPAR_TILE * GPAR_TILE;
  const size_t MM = (size_t)M * M;
  double* G = stats[0]; double* H = G + MM; double* gh = H + MM; double* dyy = gh + 2 * Mpad;
  double yy = 0.0;
  GCU(cudaMemcpyAsync(&yy, dyy, sizeof(double), cudaMemcpyDeviceToHost, c0->stream));
  for (int i = 0; i < n; i++) { GCU(cudaSetDevice(g->dev[i])); GCU(cudaStreamSynchronize(g->ctx[i]->stream)); }
  GCU(cudaSetDevice(g->dev[0]));
  if (ill && want_grad) rc = dtc_tail_whitened(c0, p, vfe, jitter, Ntot, G, H, gh, gh + Mpad, yy, val, grad, nullptr);
  else rc = dtc_tail(c0, kernel, p, vfe, jitter, Ntot, G, H, gh, gh + Mpad, yy, val, grad, nullptr, ill);
  if (rc != GPAR_OK) return group_fail(g, rc, "member 0: %s", gpar_last_error(c0));
  return GPAR_OK;
}

// ONE scaled-GPAR objective (gpar_scaled_dtc; dtc.jl:83-128) whose rows are sharded over the members: member i
// holds the FULL (t, y) — 16 bytes per step, the 1 x N filter is cheap and every member runs it — the same pseudo-inputs, and
// rows [row_lo[i], row_lo[i] + N_i) of the inputs X (consecutive slices, first rows multiples of 4).  The N x M work (kernel
// panel, whitening passes, SYRK) is sliced; the filter carry of the M whitened columns crosses the slice boundaries through
// ONE all-gather of the slice summaries (D x D transition product + D x M exit state each), and ONE all-reduce sums
// (G, g) before member 0 runs the M x M tail: 8 (M^2 + M + n (D^2 + D M)) bytes per evaluation on NVLink.
// grad (nullable): the five raw-parameter derivatives (gpar_scaled_dtc_grad) — one more all-gather (three tangent summaries);
// poorly conditioned cov(u): whitened coordinates on every slice, as gpar_scaled_dtc_grad does on one device.
int gpar_group_scaled_dtc_sharded(gpar_group* g, int k_time, int k_out, const double theta[5], const int64_t* row_lo, double* val, double* grad) {
  if (!g) return GPAR_ERR_INVALID;
  if (!theta || !row_lo || !val) return group_fail(g, GPAR_ERR_INVALID, "scaled_dtc_sharded: theta, row_lo and val must not be NULL");
  const int n = (int)g->ctx.size();
  const bool want_grad = grad != nullptr;
  gpar_ctx* c0 = g->ctx[0];
  int64_t expect = 0;
  for (int i = 0; i < n; i++) {
    gpar_ctx* c = g->ctx[i];
    if (c->M != c0->M || c->Dz != c0->Dz || c->Nt != c0->Nt)
      return group_fail(g, GPAR_ERR_INVALID, "scaled_dtc_sharded: member %d holds %lld pseudo-inputs (dimension %d) and %lld times, member 0 %lld (%d) and %lld",
                        i, (long long)c->M, c->Dz, (long long)c->Nt, (long long)c0->M, c0->Dz, (long long)c0->Nt);
    if (row_lo[i] != expect) return group_fail(g, GPAR_ERR_INVALID, "scaled_dtc_sharded: member %d starts at row %lld, the slices before it end at %lld", i, (long long)row_lo[i], (long long)expect);
    expect += c->N;
  }
  if (expect != c0->Nt) return group_fail(g, GPAR_ERR_INVALID, "scaled_dtc_sharded: the slices cover %lld rows, the sequence has %lld", (long long)expect, (long long)c0->Nt);
  std::vector<int> st;
  auto members_ok = [&]() -> int {
    for (int i = 0; i < n; i++)
      if (st[i] != GPAR_OK) return group_fail(g, st[i], "member %d (device %d): %s", i, g->dev[i], gpar_last_error(g->ctx[i]));
    return GPAR_OK;
  };
  run_members(g, st, [&](int i) { return scaled_slice_phase1(g->ctx[i], k_time, k_out, theta, row_lo[i], want_grad); });
  int rc = members_ok();
  if (rc != GPAR_OK) return rc;
  const size_t sc = c0->slice.summary_count;
  std::vector<const double*> sp(n); std::vector<double*> rp(n), stats(n);
  for (int i = 0; i < n; i++) {
    GCU(cudaSetDevice(g->dev[i]));
    GCU(g->recv[i].reserve((want_grad ? 3 : 1) * sc * n * sizeof(double)));
    sp[i] = g->ctx[i]->slice.summary; rp[i] = g->recv[i].as<double>(); stats[i] = g->ctx[i]->slice.G;
  }
  rc = group_allgather(g, sp, rp, sc);
  if (rc != GPAR_OK) return rc;
  run_members(g, st, [&](int i) { return scaled_slice_phase2(g->ctx[i], rp[i], i); });
  rc = members_ok();
  if (rc != GPAR_OK) return rc;
  rc = group_allreduce_sum(g, stats, c0->slice.stats_count);
  if (rc != GPAR_OK) return rc;
  for (int i = 0; i < n; i++) { GCU(cudaSetDevice(g->dev[i])); GCU(cudaStreamSynchronize(g->ctx[i]->stream)); }
  if (!want_grad) {
    GCU(cudaSetDevice(g->dev[0]));
    rc = scaled_slice_finish(c0, val);
    if (rc != GPAR_OK) return group_fail(g, rc, "member 0: %s", gpar_last_error(c0));
    return GPAR_OK;
  }
  // gradient: every member runs the M x M tail on the summed statistics (P, w: no broadcast needed) and the zero-start tangent
  // responses of its chunks; a SECOND all-gather carries the three tangent summaries over the slice boundaries; the five
  // partial sums of the members are added on the host
  run_members(g, st, [&](int i) { return scaled_slice_grad_phase3(g->ctx[i]); });
  rc = members_ok();
  if (rc != GPAR_OK) return rc;
  for (int i = 0; i < n; i++) sp[i] = g->ctx[i]->slice.summary2;
  rc = group_allgather(g, sp, rp, 3 * sc);
  if (rc != GPAR_OK) return rc;
  std::vector<double> s5((size_t)n * 5, 0.0);
  run_members(g, st, [&](int i) { return scaled_slice_grad_phase4(g->ctx[i], rp[i], i, s5.data() + (size_t)i * 5); });
  rc = members_ok();
  if (rc != GPAR_OK) return rc;
  double tot[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < n; i++) for (int q = 0; q < 5; q++) tot[q] += s5[(size_t)i * 5 + q];
  rc = scaled_slice_grad_finish(c0, tot, val, grad);
  if (rc != GPAR_OK) return group_fail(g, rc, "member 0: %s", gpar_last_error(c0));
  return GPAR_OK;
}

// n doubles from member src — `host` if given, else the resident result of its last gpar_lgssm_smooth /
// gpar_scaled_predict (posterior means) — to EVERY member's chain buffer (ncclBroadcast); out (nullable): host copy.
int gpar_group_broadcast(gpar_group* g, int32_t src, const double* host, int64_t n, double* out) {
  if (!g) return GPAR_ERR_INVALID;
  const int nm = (int)g->ctx.size();
  if (src < 0 || src >= nm || n < 1) return group_fail(g, GPAR_ERR_INVALID, "group_broadcast: bad source member %d or length %lld", src, (long long)n);
  gpar_ctx* sc = g->ctx[src];
  // a merged train+test problem (gpar_set_merged): the resident result is in SORTED train+test order — what goes down the
  // chain is its gather at the N* test locations in test order (gpar_take_test's first array), n = N*
  const bool merged = !host && sc->merged_Ns > 0 && sc->res_a && sc->res_len == sc->merged_N + sc->merged_Ns;
  if (merged && n != sc->merged_Ns)
    return group_fail(g, GPAR_ERR_INVALID, "group_broadcast: member %d holds a merged result with %lld test locations, asked for %lld values",
                      src, (long long)sc->merged_Ns, (long long)n);
  if (!host && !merged && (!sc->res_a || sc->res_len < n)) return group_fail(g, GPAR_ERR_INVALID, "group_broadcast: member %d holds no resident result of length >= %lld", src, (long long)n);
  for (int i = 0; i < nm; i++) {
    GCU(cudaSetDevice(g->dev[i]));
    GCU(g->ctx[i]->chain.reserve((size_t)n * sizeof(double)));
    g->ctx[i]->chain_n = n;
  }
  GCU(cudaSetDevice(g->dev[src]));
  if (host) GCU(cudaMemcpyAsync(sc->chain.p, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, sc->stream));
  else if (merged) { const int rc = merged_gather_test(sc, sc->chain.as<double>(), nullptr); if (rc != GPAR_OK) return group_fail(g, rc, "member %d: %s", src, gpar_last_error(sc)); }
  else GCU(cudaMemcpyAsync(sc->chain.p, sc->res_a, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, sc->stream));
  if (g->loopback) {
    GCU(cudaStreamSynchronize(sc->stream));
    for (int i = 0; i < nm; i++)
      if (i != src) GCU(cudaMemcpyAsync(g->ctx[i]->chain.p, sc->chain.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, g->ctx[i]->stream));
  } else {
    GNC(g->GroupStart());
    for (int i = 0; i < nm; i++) GNC(g->Broadcast(g->ctx[i]->chain.p, g->ctx[i]->chain.p, (size_t)n, ncclDouble, src, g->comm[i], g->ctx[i]->stream));
    GNC(g->GroupEnd());
  }
  for (int i = 0; i < nm; i++) {
    GCU(cudaSetDevice(g->dev[i]));
    if (out && i == (src + 1) % nm) GCU(cudaMemcpyAsync(out, g->ctx[i]->chain.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, g->ctx[i]->stream));   // a RECEIVER's copy
    GCU(cudaStreamSynchronize(g->ctx[i]->stream));
  }
  return GPAR_OK;
}

// Whole fits of a task list: every member takes the next task (longest first) until none is left; per task the
// data go to the member's device once and the optimiser — Nelder-Mead (the reference's, dtc.jl:58-61) or L-BFGS on
// the analytic gradients — runs `iterations` iterations on the blocking objective (dtc.jl:29-61: minimise -dtc;
// temporal_gp_inference.jl:69-82: minimise -logpdf for a time-only task, D = 0).
// minimum[k], minimizer[5 k .. 5 k + 4] (NaN-padded for 3-parameter tasks), f_calls[k], member_of[k] (nullable).
int gpar_group_fit(gpar_group* g, const double* t, int64_t N, const gpar_fit_task* tasks, int32_t ntasks, int k_time, int k_out,
                   int32_t optimizer, int32_t iterations, double* minimum, double* minimizer, int32_t* f_calls, int32_t* member_of) {
  if (!g) return GPAR_ERR_INVALID;
  if (!t || N < 1 || !tasks || ntasks < 1 || !minimum || !minimizer) return group_fail(g, GPAR_ERR_INVALID, "group_fit: t, tasks, minimum and minimizer must be given");
  if (optimizer != GPAR_OPT_NELDER_MEAD && optimizer != GPAR_OPT_LBFGS) return group_fail(g, GPAR_ERR_INVALID, "group_fit: unknown optimizer %d", optimizer);
  for (int k = 0; k < ntasks; k++)
    if (!tasks[k].y || (tasks[k].D > 0 && (!tasks[k].X || !tasks[k].Z || tasks[k].M < 1)) || tasks[k].D < 0)
      return group_fail(g, GPAR_ERR_INVALID, "group_fit: task %d is incomplete", k);
  const int n = (int)g->ctx.size();
  std::vector<int> order(ntasks);
  std::iota(order.begin(), order.end(), 0);
  auto cost = [&](int k) { return tasks[k].D == 0 ? 0.02 : (1.0 + 0.03 * tasks[k].D) * (double)tasks[k].M * (double)tasks[k].M; };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
  std::atomic<int> next(0);
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<double> rows((size_t)n * ntasks * 8, nan);      // per member: its tasks' (task, minimum, minimiser[5], calls)
  std::vector<int> st;
  run_members(g, st, [&](int i) {
    gpar_ctx* c = g->ctx[i];
    int rc = gpar_set_times(c, t, N);
    if (rc != GPAR_OK) return rc;
    for (;;) {
      const int q = next.fetch_add(1);
      if (q >= ntasks) break;
      const int k = order[q];
      const gpar_fit_task& tk = tasks[k];
      rc = gpar_set_noise_vector(c, nullptr, 0);
      if (rc == GPAR_OK) rc = gpar_set_outputs(c, tk.y, N, 1);
      if (rc == GPAR_OK && tk.D > 0) rc = gpar_set_inputs(c, tk.X, tk.D, N);
      if (rc == GPAR_OK && tk.D > 0) rc = gpar_set_pseudo(c, tk.Z, tk.D, tk.M);
      if (rc != GPAR_OK) return rc;
      const int np = tk.D > 0 ? 5 : 3;
      int hard = GPAR_OK;
      auto f = [&](const double* th) -> double {
        double v = 0.0;
        int r = tk.D > 0 ? gpar_scaled_dtc(c, k_time, k_out, th, &v, nullptr) : gpar_lgssm_logpdf(c, k_time, th, 1, &v);
        if (r == GPAR_ERR_NOT_POSDEF) return std::numeric_limits<double>::infinity();      // what the Python mirror does with a PosDefException
        if (r != GPAR_OK) { hard = r; return std::numeric_limits<double>::infinity(); }
        return -v;
      };
      auto fgr = [&](const double* th, double* gr) -> double {      // value and gradient of the NEGATED objective
        double v = 0.0;
        int r = tk.D > 0 ? gpar_scaled_dtc_grad(c, k_time, k_out, th, &v, gr) : gpar_lgssm_logpdf_grad(c, k_time, th, 1, &v, gr);
        if (r != GPAR_OK) { if (r != GPAR_ERR_NOT_POSDEF) hard = r; for (int j = 0; j < np; j++) gr[j] = 0.0; return std::numeric_limits<double>::infinity(); }
        for (int j = 0; j < np; j++) gr[j] = -gr[j];
        return -v;
      };
      double xb[5] = {nan, nan, nan, nan, nan}, fb = nan; int calls = 0;
      if (optimizer == GPAR_OPT_LBFGS) lbfgs(fgr, tk.theta0, np, iterations, 1e-6, 1e-10, xb, &fb, &calls);
      else nelder_mead(f, tk.theta0, np, iterations, 1e-8, xb, &fb, &calls);
      if (hard != GPAR_OK) return hard;
      double* row = rows.data() + ((size_t)i * ntasks + k) * 8;
      row[0] = k; row[1] = fb; for (int j = 0; j < 5; j++) row[2 + j] = xb[j]; row[7] = calls;
    }
    return (int)GPAR_OK;
  });
  for (int i = 0; i < n; i++)
    if (st[i] != GPAR_OK) return group_fail(g, st[i], "group_fit: member %d (device %d): %s", i, g->dev[i], gpar_last_error(g->ctx[i]));
  std::vector<double> table((size_t)n * ntasks * 8);
  int rc = allgather_rows(g, rows.data(), ntasks * 8, table.data());
  if (rc != GPAR_OK) return rc;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < ntasks; k++) {
      const double* row = table.data() + ((size_t)i * ntasks + k) * 8;
      if (row[0] != row[0]) continue;       // not this member's task
      minimum[k] = row[1];
      for (int j = 0; j < 5; j++) minimizer[(size_t)5 * k + j] = row[2 + j];
      if (f_calls) f_calls[k] = (int32_t)row[7];
      if (member_of) member_of[k] = i;
    }
  return GPAR_OK;
}

// compute_q_u (gpar_scaled_inference.jl:141-196) with the rows sharded like gpar_group_scaled_dtc_sharded: the same two collectives,
// bare Cuu (usually poorly conditioned: every member then whitens its panel by L_u), the M x M tail on member 0.
// m_e[M], Dinv[M x M], U_u[M x M] column-major, as gpar_compute_q_u.  With it a single output whose N x M panel exceeds one device
// can be fitted (gpar_group_fit_sharded) AND predicted: the draws from q(u) and gpar_scaled_predict need no N x M array.
static int group_qu_stats(gpar_group* g, int k_time, int k_out, const double params[5], const int64_t* row_lo, const char* who) {
  const int n = (int)g->ctx.size();
  gpar_ctx* c0 = g->ctx[0];
  int64_t expect = 0;
  for (int i = 0; i < n; i++) {
    gpar_ctx* c = g->ctx[i];
    if (c->M != c0->M || c->Dz != c0->Dz || c->Nt != c0->Nt) return group_fail(g, GPAR_ERR_INVALID, "%s: member %d holds other pseudo-inputs or times than member 0", who, i);
    if (row_lo[i] != expect) return group_fail(g, GPAR_ERR_INVALID, "%s: member %d starts at row %lld, the slices before it end at %lld", who, i, (long long)row_lo[i], (long long)expect);
    expect += c->N;
  }
  if (expect != c0->Nt) return group_fail(g, GPAR_ERR_INVALID, "%s: the slices cover %lld rows, the sequence has %lld", who, (long long)expect, (long long)c0->Nt);
  std::vector<int> st;
  auto members_ok = [&]() -> int {
    for (int i = 0; i < n; i++)
      if (st[i] != GPAR_OK) return group_fail(g, st[i], "member %d (device %d): %s", i, g->dev[i], gpar_last_error(g->ctx[i]));
    return GPAR_OK;
  };
  run_members(g, st, [&](int i) { return scaled_slice_phase1(g->ctx[i], k_time, k_out, params, row_lo[i], false, true); });
  int rc = members_ok();
  if (rc != GPAR_OK) return rc;
  const size_t sc = c0->slice.summary_count;
  std::vector<const double*> sp(n); std::vector<double*> rp(n), stats(n);
  for (int i = 0; i < n; i++) {
    GCU(cudaSetDevice(g->dev[i]));
    GCU(g->recv[i].reserve(sc * n * sizeof(double)));
    sp[i] = g->ctx[i]->slice.summary; rp[i] = g->recv[i].as<double>(); stats[i] = g->ctx[i]->slice.G;
  }
  rc = group_allgather(g, sp, rp, sc);
  if (rc != GPAR_OK) return rc;
  run_members(g, st, [&](int i) { return scaled_slice_phase2(g->ctx[i], rp[i], i); });
  rc = members_ok();
  if (rc != GPAR_OK) return rc;
  rc = group_allreduce_sum(g, stats, c0->slice.stats_count);
  if (rc != GPAR_OK) return rc;
  for (int i = 0; i < n; i++) { GCU(cudaSetDevice(g->dev[i])); GCU(cudaStreamSynchronize(g->ctx[i]->stream)); }
  GCU(cudaSetDevice(g->dev[0]));
  return GPAR_OK;
}
int gpar_group_compute_q_u_sharded(gpar_group* g, int k_time, int k_out, const double params[5], const int64_t* row_lo, double* m_e, double* Dinv, double* U_u) {
  if (!g) return GPAR_ERR_INVALID;
  if (!params || !row_lo || !m_e || !Dinv || !U_u) return group_fail(g, GPAR_ERR_INVALID, "compute_q_u_sharded: NULL argument");
  int rc = group_qu_stats(g, k_time, k_out, params, row_lo, "compute_q_u_sharded");
  if (rc != GPAR_OK) return rc;
  rc = scaled_slice_qu_finish(g->ctx[0], m_e, Dinv, U_u);
  if (rc != GPAR_OK) return group_fail(g, rc, "member 0: %s", gpar_last_error(g->ctx[0]));
  return GPAR_OK;
}
// S seeded draws from that q(u) on member 0's device (gpar_sample_q_u's Philox sampler) -> W = U_u \ eps (M x S, host; also resident on
// member 0 for a gpar_scaled_predict(W = NULL) there), eps_out (nullable).  No host numerics between the sharded statistics and W.
int gpar_group_sample_q_u_sharded(gpar_group* g, int k_time, int k_out, const double params[5], const int64_t* row_lo, uint64_t seed, int32_t S,
                                  double* W_out, double* eps_out) {
  if (!g) return GPAR_ERR_INVALID;
  if (!params || !row_lo || S < 1) return group_fail(g, GPAR_ERR_INVALID, "sample_q_u_sharded: params and row_lo must be given, S >= 1");
  int rc = group_qu_stats(g, k_time, k_out, params, row_lo, "sample_q_u_sharded");
  if (rc != GPAR_OK) return rc;
  rc = scaled_slice_qu_sample(g->ctx[0], seed, S, W_out, eps_out);
  if (rc != GPAR_OK) return group_fail(g, rc, "member 0: %s", gpar_last_error(g->ctx[0]));
  return GPAR_OK;
}

// ONE fit with every device working on every evaluation: the rows of the objective are sharded over the members
// (gpar_group_scaled_dtc_sharded on the slices already resident: full (t, y), Z and the member's rows of X — e.g. loaded once by
// the caller) and the host optimiser of gpar_group_fit — Nelder-Mead (dtc.jl:58-61) or L-BFGS on the sharded gradient — drives
// it.  For a single output too large for one device's memory or time budget.  minimizer[5], minimum (of the NEGATED objective).
int gpar_group_fit_sharded(gpar_group* g, int k_time, int k_out, const int64_t* row_lo, const double theta0[5], int32_t optimizer, int32_t iterations,
                           double* minimum, double* minimizer, int32_t* f_calls) {
  if (!g) return GPAR_ERR_INVALID;
  if (!row_lo || !theta0 || !minimum || !minimizer) return group_fail(g, GPAR_ERR_INVALID, "group_fit_sharded: row_lo, theta0, minimum and minimizer must be given");
  if (optimizer != GPAR_OPT_NELDER_MEAD && optimizer != GPAR_OPT_LBFGS) return group_fail(g, GPAR_ERR_INVALID, "group_fit_sharded: unknown optimizer %d", optimizer);
  int hard = GPAR_OK;
  const double inf = std::numeric_limits<double>::infinity();
  auto f = [&](const double* th) -> double {
    double v = 0.0;
    const int r = gpar_group_scaled_dtc_sharded(g, k_time, k_out, th, row_lo, &v, nullptr);
    if (r == GPAR_ERR_NOT_POSDEF) return inf;
    if (r != GPAR_OK) { if (hard == GPAR_OK) hard = r; return inf; }
    return -v;
  };
  auto fgr = [&](const double* th, double* gr) -> double {
    double v = 0.0;
    const int r = gpar_group_scaled_dtc_sharded(g, k_time, k_out, th, row_lo, &v, gr);
    if (r != GPAR_OK) { if (r != GPAR_ERR_NOT_POSDEF && hard == GPAR_OK) hard = r; for (int j = 0; j < 5; j++) gr[j] = 0.0; return inf; }
    for (int j = 0; j < 5; j++) gr[j] = -gr[j];
    return -v;
  };
  double xb[5], fb = std::numeric_limits<double>::quiet_NaN(); int calls = 0;
  for (int j = 0; j < 5; j++) xb[j] = fb;
  if (optimizer == GPAR_OPT_LBFGS) lbfgs(fgr, theta0, 5, iterations, 1e-6, 1e-10, xb, &fb, &calls);
  else nelder_mead(f, theta0, 5, iterations, 1e-8, xb, &fb, &calls);
  if (hard != GPAR_OK) return hard;          // g->err holds the message of the failing evaluation (or a later one of the same kind)
  *minimum = fb;
  for (int j = 0; j < 5; j++) minimizer[j] = xb[j];
  if (f_calls) *f_calls = calls;
  return GPAR_OK;
}

// Column d of the resident inputs X (N records of D doubles) <- col (host), or, with col == NULL, the context's
// chain buffer (filled by gpar_group_broadcast): the predicted means of an earlier output become an input feature
// of the later ones without a host round trip (GPAR_scaled_examples.jl:172).
__global__ void set_column_kernel(double* __restrict__ X, int D, int d, const double* __restrict__ col, int64_t N) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) X[i * D + d] = col[i];
}

int gpar_set_inputs_column(gpar_ctx* ctx, int32_t d, const double* col) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (ctx->N < 1 || d < 0 || d >= ctx->D) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_inputs_column: column %d outside the resident inputs (D = %d, N = %lld)", d, ctx->D, (long long)ctx->N);
  CU(cudaSetDevice(ctx->device));
  if (col) {
    CU(ctx->chain.reserve((size_t)ctx->N * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->chain.p, col, (size_t)ctx->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->chain_n = ctx->N;
  } else if (ctx->chain_n != ctx->N) {
    return gpar_fail(ctx, GPAR_ERR_INVALID, "set_inputs_column: chain buffer holds %lld values, the inputs have %lld records", (long long)ctx->chain_n, (long long)ctx->N);
  }
  LAUNCH(ctx, set_column_kernel, (unsigned)((ctx->N + 255) / 256), 256, 0, ctx->X.as<double>(), ctx->D, d, ctx->chain.as<double>(), ctx->N);
  CU(cudaStreamSynchronize(ctx->stream));
  return GPAR_OK;
}

}  // extern "C"
