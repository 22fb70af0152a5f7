// b200md -- context lifecycle, host<->device staging of atoms and neighbor lists,
// device prefix scan, options/counters.  No CPU fallback anywhere: without a usable
// sm_100a device every call fails with B200MD_ERR_CUDA.

#include "common.cuh"

#include <mutex>

static std::string g_create_error;
static std::mutex g_mutex;

extern "C" int b200md_version(void) { return B200MD_VERSION; }

extern "C" const char *b200md_last_error(const b200md_ctx *ctx)
{
  if (ctx) return ctx->err.c_str();
  return g_create_error.c_str();
}

extern "C" int b200md_device_count(void)
{
  int ndev = 0, usable = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  for (int d = 0; d < ndev; d++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) usable++;
  }
  return usable == ndev ? ndev : 0;    // a mixed box: say 0 and let the caller name the device explicitly
}

extern "C" int b200md_create(int device, b200md_ctx **out)
{
  std::lock_guard<std::mutex> lk(g_mutex);
  if (!out) {
    g_create_error = "b200md_create: out is NULL";
    return B200MD_ERR_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("b200md_create: no CUDA device available (") +
        (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
        "); this library has no CPU fallback";
    return B200MD_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "b200md_create: device index out of range";
    return B200MD_ERR_ARG;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return B200MD_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "b200md_create: device " + std::string(prop.name) + " is sm_" +
        std::to_string(prop.major) + std::to_string(prop.minor) +
        "; this library carries sm_100a code only";
    return B200MD_ERR_CUDA;
  }
  b200md_ctx *c = new b200md_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete c;
    return B200MD_ERR_CUDA;
  }
  if (c->scal.reserve(64) != cudaSuccess || c->flags.reserve(16 + B200MD_MAX_D2H_CHUNKS) != cudaSuccess ||
      c->pin_scal.reserve(64) != cudaSuccess) {
    g_create_error = "b200md_create: device allocation failed";
    delete c;
    return B200MD_ERR_CUDA;
  }
  cudaMemsetAsync(c->scal.p, 0, 64 * sizeof(double), c->stream);
  cudaMemsetAsync(c->flags.p, 0, (16 + B200MD_MAX_D2H_CHUNKS) * sizeof(int), c->stream);
  cudaStreamSynchronize(c->stream);
  *out = c;
  return B200MD_OK;
}

void b200md_system_free(b200md_ctx *c);    // system.cu

extern "C" void b200md_destroy(b200md_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  b200md_system_free(c);
  c->x_aos.release(); c->xq.release(); c->f.release(); c->type.release(); c->tag.release();
  c->pin_f.release(); c->pin_scal.release(); c->pin_pa.release(); c->eatom_d.release(); c->vatom_d.release(); c->scal.release(); c->flags.release();
  c->list_off.release(); c->list_num.release(); c->list_val.release(); c->xhold.release();
  c->map_d.release(); c->short_idx.release(); c->short_num.release();
  c->lj_off.release(); c->lj_num.release(); c->lj_val.release(); c->ljp_ab.release();
  c->short_idx_t.release(); c->short_num_t.release(); c->lj_val_t.release(); c->lj_num_t.release(); c->xhold_t.release();
  c->ljp_tmp.release(); c->ljp_scan.release();
  c->strag_flag.release(); c->strag_list.release(); c->strag_dev.release(); c->strag_pin.release();
  if (c->ev_tight) cudaEventDestroy(c->ev_tight);
  if (c->halo_stream) cudaStreamDestroy(c->halo_stream);
  for (cudaEvent_t e : {c->ev_ready, c->ev_fwd, c->ev_reb, c->ev_rev})
    if (e) cudaEventDestroy(e);
  c->cen_list.release(); c->cen_key.release(); c->cen_scan.release(); c->nM.release(); c->nS.release(); c->det_fb.release(); c->det_j.release();
  c->spl_frho.release(); c->spl_rhor.release(); c->spl_z2r.release(); c->spl_pair.release();
  c->rho.release(); c->fp.release(); c->ea_off.release(); c->ea_num.release(); c->ea_val.release();
  c->ang_list.release(); c->ang_key.release(); c->ang_scan.release(); c->fp_tmp.release(); c->det_ffix.release(); c->det_part.release();
  c->ec_off.release(); c->ec_num.release(); c->ec_val.release(); c->ec_cap.release(); c->ec_df.release();
  c->bin_of.release(); c->bin_count.release(); c->bin_start.release(); c->bin_atoms.release();
  c->stencil_d.release(); c->scan_tmp.release(); c->scan_tmp64.release();
  for (int k = 0; k < 8; k++)
    if (c->ev[k]) cudaEventDestroy(c->ev[k]);
  for (TimedLaunch &t : c->pending) {
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : c->copy_ev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->copy_done)
    if (e) cudaEventDestroy(e);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (cudaEvent_t e : c->up_ev)
    if (e) cudaEventDestroy(e);
  if (c->up_stream) cudaStreamDestroy(c->up_stream);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int b200md_set_option(b200md_ctx *c, const char *name, long long value)
{
  if (!c || !name) return B200MD_ERR_ARG;
  std::string n(name);
  if (n == "deterministic") {
    c->deterministic = value ? 1 : 0;
    if (c->deterministic && !c->det_part.p) {    // table of per-block partial sums of the global accumulators
      cudaSetDevice(c->device);
      if (c->det_part.reserve((size_t) B200MD_DET_BLOCKS * 16) != cudaSuccess) {
        c->fail("deterministic mode: cannot allocate the partial-sum table");
        c->deterministic = 0;
        return B200MD_ERR_CUDA;
      }
      cudaMemsetAsync(c->det_part.p, 0, (size_t) B200MD_DET_BLOCKS * 16 * sizeof(double), c->stream);
    }
  }
  else if (n == "margin") {
    c->margin_opt = 1.0e-3 * (double) value;
    c->inner_valid = false;
  } else if (n == "margin_tight") {
    c->margin_t_opt = 1.0e-3 * (double) value;
    c->inner_valid = false;
  } else if (n == "sync_timing") c->sync_timing = value ? 1 : 0;
  else if (n == "f_overwrite") c->f_overwrite = value ? 1 : 0;
  else if (n == "peratom") c->peratom_opt = value ? 1 : 0;
  else if (n == "fp_gated") c->fp_gated = value ? 1 : 0;
  else if (n == "p2p_halo") c->p2p_halo = value ? 1 : 0;
  else if (n == "lj_pairs") {
    c->lj_pairs = value ? 1 : 0;
    c->inner_valid = false;
  } else if (n == "aeam_cluster") {
    c->aeam_cluster = (int) (value < 0 ? 0 : (value > 2 ? 2 : value));
    c->inner_valid = false;
  } else if (n == "overlap_halo") {
    c->overlap_halo = (int) (value < 0 ? 0 : value);
    c->inner_valid = false;
  } else if (n == "one_pass_neigh") c->one_pass_neigh = value ? 1 : 0;
  else if (n == "fuse_integrate") c->fuse_integrate = value ? 1 : 0;
  else if (n == "neigh_unroll") c->neigh_unroll = (int) value;
  else if (n == "peer_vote") c->peer_vote = value ? 1 : 0;
  else if (n == "flat_halo") c->flat_halo = value ? 1 : 0;
  else if (n == "split_elems") c->split_elems = value == 2 ? 2 : 3;
  else if (n == "aeam_variant") c->aeam_variant = (int) value;
  else if (n == "force_rebuild") c->force_rebuild = value ? 1 : 0;
  else if (n == "aeam_sort_rows") {
    c->aeam_sort_rows = value ? 1 : 0;
    c->inner_valid = false;
  } else if (n == "ang_ctas") c->ang_ctas = (int) (value < 1 ? 1 : value);
  else if (n == "h2d_chunks") {
    c->h2d_chunks = (int) (value < 1 ? 1 : (value > B200MD_MAX_D2H_CHUNKS ? B200MD_MAX_D2H_CHUNKS : value));
    c->inner_valid = false;
  } else if (n == "h2d_ramp") {
    c->h2d_ramp = (int) value;
    c->inner_valid = false;
  } else if (n == "d2h_min_atoms") c->d2h_min_atoms = (int) value;
  else if (n == "d2h_chunks") c->d2h_chunks = (int) (value < 1 ? 1 : (value > B200MD_MAX_D2H_CHUNKS ? B200MD_MAX_D2H_CHUNKS : value));
  else {
    c->fail("unknown option " + n);
    return B200MD_ERR_ARG;
  }
  return B200MD_OK;
}

// on-demand sums of per-row counts (measurement queries, never on the step path)
__global__ void count_sum_kernel(const int *__restrict__ v, long long n, unsigned long long *__restrict__ out)
{
  unsigned long long s = 0;
  for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    s += (unsigned long long) v[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}
static long long device_sum(b200md_ctx *c, const int *v, long long n)
{
  if (!v || n <= 0) return 0;
  cudaSetDevice(c->device);
  unsigned long long *acc = (unsigned long long *) (c->scal.p + 60);
  unsigned long long h = 0;
  cudaMemsetAsync(acc, 0, sizeof(h), c->stream);
  count_sum_kernel<<<c->num_sms * 2, 256, 0, c->stream>>>(v, n, acc);
  cudaMemcpyAsync(&h, acc, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
  return (long long) h;
}

extern "C" long long b200md_get_counter(b200md_ctx *c, const char *name)
{
  if (!c || !name) return -1;
  std::string n(name);
  // entries of the rows the force kernels stream right now (summed on the device when asked)
  if (n == "lj_entries_tight")
    return (c->tight_valid && c->lj_pairs) ? device_sum(c, c->lj_num_t.p, 4LL * c->ljp_P) : -1;
  if (n == "short_entries_tight") return c->tight_valid ? device_sum(c, c->short_num_t.p, c->list_inum) : -1;
  if (n == "short_entries_owned") return c->inner_valid && c->rebomos_ready ? device_sum(c, c->short_num.p, c->list_inum) : -1;
  if (n == "aeam_entries")
    return (c->inner_valid && c->aeam_ready)
               ? (c->aeam_cluster == 1 ? device_sum(c, c->ec_num.p, (c->list_inum + 3) / 4) : device_sum(c, c->ea_num.p, c->list_inum))
               : -1;
  if (n == "upload_stragglers") return c->n_strag;
  if (n == "master_entries") return c->list_valid ? (long long) c->list_entries_used : -1;
  if (n == "kernel_launches") return c->n_launch;
  if (n == "list_uploads") return c->n_list_upload;
  if (n == "compute_calls") return c->n_compute;
  if (n == "inner_rebuilds") return c->n_inner_rebuild;
  if (n == "h2d_bytes") return c->h2d_bytes;
  if (n == "d2h_bytes") return c->d2h_bytes;
  if (n == "lj_entries") return c->n_lj_entries;
  if (n == "short_entries") return c->n_short_entries;
  if (n == "num_sms") return c->num_sms;
  if (n == "p2p_exchanges") return c->n_p2p;
  if (n == "tight_refreshes") return c->n_tight;
  if (n == "pipelined_calls") return c->n_pipelined;
  if (n == "pipelined_redos") return c->n_redo;
  return -1;
}

extern "C" double b200md_last_kernel_ms(b200md_ctx *c, const char *name)
{
  if (!c || !name) return -1.0;
  auto it = c->kstat.find(name);
  if (it == c->kstat.end()) return -1.0;
  return it->second.last_ms;
}

extern "C" int b200md_kernel_stats(b200md_ctx *c, int index, char *name_out, int name_cap, double *total_ms,
                                   long long *count)
{
  if (!c) return B200MD_ERR_ARG;
  if (index < 0 || index >= (int) c->kstat.size()) return 1;
  auto it = c->kstat.begin();
  std::advance(it, index);
  if (name_out && name_cap > 0) {
    strncpy(name_out, it->first.c_str(), name_cap - 1);
    name_out[name_cap - 1] = '\0';
  }
  if (total_ms) *total_ms = it->second.total_ms;
  if (count) *count = it->second.count;
  return B200MD_OK;
}

extern "C" int b200md_kernel_stats_reset(b200md_ctx *c)
{
  if (!c) return B200MD_ERR_ARG;
  c->kstat.clear();
  return B200MD_OK;
}

extern "C" void *b200md_stream(b200md_ctx *c) { return c ? (void *) c->stream : nullptr; }

extern "C" int b200md_event_record(b200md_ctx *c, int slot)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, slot >= 0 && slot < 8, "event_record: slot must be 0..7");
  CUDA_TRY(c, cudaSetDevice(c->device));
  if (!c->ev[slot]) CUDA_TRY(c, cudaEventCreate(&c->ev[slot]));
  CUDA_TRY(c, cudaEventRecord(c->ev[slot], c->stream));
  return B200MD_OK;
}

extern "C" double b200md_event_elapsed_ms(b200md_ctx *c, int a, int b)
{
  if (!c || a < 0 || a >= 8 || b < 0 || b >= 8 || !c->ev[a] || !c->ev[b]) return -1.0;
  if (cudaEventSynchronize(c->ev[b]) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) return -1.0;
  return (double) ms;
}

extern "C" int b200md_host_register(void *p, size_t bytes)
{
  if (!p || !bytes) return B200MD_ERR_ARG;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
  if (e == cudaSuccess) return B200MD_OK;
  cudaGetLastError();    // not sticky: the caller carries on with pageable memory
  return (e == cudaErrorHostMemoryAlreadyRegistered) ? B200MD_OK : B200MD_ERR_CUDA;
}

extern "C" int b200md_host_unregister(void *p)
{
  if (!p) return B200MD_ERR_ARG;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) cudaGetLastError();
  return e == cudaSuccess ? B200MD_OK : B200MD_ERR_CUDA;
}

extern "C" void *b200md_host_alloc(size_t bytes)
{
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void b200md_host_free(void *p)
{
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------- roofline denominators measured in place
// DFMA-saturating kernel: 8 independent chains per thread (MEASURED_PEAKS.json carries HBM and bf16 numbers
// only; the FP64 roofline fraction needs its own denominator on the same box, same clocks)
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b)
{
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int k = 0; k < iters; k++) {
    v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
    v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
  }
  const double s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
  if (s == 123.456) out[0] = s;    // never true; keeps the chains alive
}
__global__ void __launch_bounds__(256) copy_peak_kernel(const double4 *__restrict__ in, double4 *__restrict__ out, size_t n)
{
  for (size_t i = blockIdx.x * (size_t) 256 + threadIdx.x; i < n; i += (size_t) gridDim.x * 256) out[i] = in[i];
}

extern "C" int b200md_measure_peaks(b200md_ctx *c, double *fp64_tflops, double *hbm_gbs)
{
  if (!c) return B200MD_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaEvent_t e0, e1;
  CUDA_TRY(c, cudaEventCreate(&e0));
  CUDA_TRY(c, cudaEventCreate(&e1));
  float ms = 0.f;
  if (fp64_tflops) {
    const int iters = 1 << 14, blocks = c->num_sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, c->stream);
      dfma_peak_kernel<<<blocks, 256, 0, c->stream>>>(c->scal.p + 48, iters, 1.0000001, 1.0e-9);
      cudaEventRecord(e1, c->stream);
      CUDA_TRY(c, cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    *fp64_tflops = best;
  }
  if (hbm_gbs) {
    const size_t n = (size_t) 1 << 25;    // 2 x 1 GiB of double4: far beyond the 126 MB L2
    DevBuf<double4> a, b;
    CUDA_TRY(c, a.reserve(n));
    CUDA_TRY(c, b.reserve(n));
    CUDA_TRY(c, cudaMemsetAsync(a.p, 0, n * sizeof(double4), c->stream));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, c->stream);
      copy_peak_kernel<<<c->num_sms * 16, 256, 0, c->stream>>>(a.p, b.p, n);
      cudaEventRecord(e1, c->stream);
      CUDA_TRY(c, cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      const double gbs = 2.0 * n * sizeof(double4) / (ms * 1e-3) / 1e9;
      if (rep > 0 && gbs > best) best = gbs;
    }
    a.release();
    b.release();
    *hbm_gbs = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// slot k = blockIdx.x: 256 threads sum contiguous chunks of the table rows in ascending order, thread 0 adds the 256
// chunk sums in ascending order -- a fixed tree -- and the rows are zeroed for the next evaluation
__global__ void __launch_bounds__(256) det_fold_kernel(double *__restrict__ part, int rows, double *__restrict__ scal)
{
  __shared__ double sh[256];
  const int k = blockIdx.x;
  const int per = (rows + 255) / 256;
  const int lo = threadIdx.x * per, hi = min(rows, lo + per);
  double s = 0.0;
  for (int b = lo; b < hi; b++) {
    const double v = part[(size_t) b * 16 + k];
    if (v != 0.0) {
      s += v;
      part[(size_t) b * 16 + k] = 0.0;
    }
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 256; w++) t += sh[w];
    scal[k] += t;
  }
}
int b200md_det_fold(b200md_ctx *c)
{
  if (!c->deterministic || !c->det_part.p) return B200MD_OK;
  det_fold_kernel<<<16, 256, 0, c->stream>>>(c->det_part.p, B200MD_DET_BLOCKS, c->scal.p);
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// forces + scalars + flags back to the host; f accumulated (default) or overwritten
int b200md_finish_compute(b200md_ctx *c, int eflag, int vflag, double *f, double *eng_vdwl, double *virial,
                          int *flags_out)
{
  const size_t n3 = 3 * (size_t) c->nall;
  int *pin_flags = (int *) (c->pin_scal.p + 32);
  int rcf = b200md_det_fold(c);
  if (rcf) return rcf;
  if (c->f_overwrite) {
    if (n3) CUDA_TRY(c, cudaMemcpyAsync(f, c->f.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  } else {
    CUDA_TRY(c, c->pin_f.reserve(n3 + 64));
    if (n3) CUDA_TRY(c, cudaMemcpyAsync(c->pin_f.p, c->f.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_TRY(c, cudaMemcpyAsync(c->pin_scal.p, c->scal.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(pin_flags, c->flags.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->d2h_bytes += (long long) (n3 * sizeof(double) + 16 * sizeof(double) + 16 * sizeof(int));
  b200md_collect_timers(c);
  memcpy(flags_out, pin_flags, 16 * sizeof(int));
  if (!c->f_overwrite) {
    const double *src = c->pin_f.p;
    for (size_t k = 0; k < n3; k++) f[k] += src[k];
  }
  if (eng_vdwl) *eng_vdwl = eflag ? c->pin_scal.p[0] : 0.0;
  if (virial)
    for (int k = 0; k < 6; k++) virial[k] = vflag ? c->pin_scal.p[1 + k] : 0.0;
  return B200MD_OK;
}

int b200md_d2h_begin(b200md_ctx *c)
{
  if (!c->copy_stream) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  c->d2h_ranges.clear();
  if (!c->f_overwrite) CUDA_TRY(c, c->pin_f.reserve(3 * (size_t) c->nall + 64));
  return B200MD_OK;
}

int b200md_d2h_range(b200md_ctx *c, int slot, double *f_host, size_t lo, size_t hi)
{
  if (hi <= lo) return B200MD_OK;
  if (slot < 0 || slot >= B200MD_MAX_D2H_CHUNKS + 2) return c->fail("d2h_range: bad slot"), B200MD_ERR_ARG;
  if (!c->copy_ev[slot]) CUDA_TRY(c, cudaEventCreateWithFlags(&c->copy_ev[slot], cudaEventDisableTiming));
  if (!c->copy_done[slot]) CUDA_TRY(c, cudaEventCreateWithFlags(&c->copy_done[slot], cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventRecord(c->copy_ev[slot], c->stream));
  CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->copy_ev[slot], 0));
  double *dst = c->f_overwrite ? f_host : c->pin_f.p;
  CUDA_TRY(c, cudaMemcpyAsync(dst + lo, c->f.p + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToHost,
                              c->copy_stream));
  CUDA_TRY(c, cudaEventRecord(c->copy_done[slot], c->copy_stream));
  c->d2h_ranges.push_back({slot, lo, hi});
  c->d2h_bytes += (long long) ((hi - lo) * sizeof(double));
  return B200MD_OK;
}

int b200md_d2h_finish(b200md_ctx *c, int eflag, int vflag, double *f_host, double *eng_vdwl, double *virial,
                      int *flags_out)
{
  int *pin_flags = (int *) (c->pin_scal.p + 32);
  int rcf = b200md_det_fold(c);
  if (rcf) return rcf;
  CUDA_TRY(c, cudaMemcpyAsync(c->pin_scal.p, c->scal.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(pin_flags, c->flags.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (!c->f_overwrite) {
    // accumulate mode: the host adds every range as soon as it has arrived, while later ranges still compute
    const double *src = c->pin_f.p;
    for (const b200md_ctx::D2HRange &r : c->d2h_ranges) {
      CUDA_TRY(c, cudaEventSynchronize(c->copy_done[r.slot]));
      for (size_t k = r.lo; k < r.hi; k++) f_host[k] += src[k];
    }
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
  c->d2h_bytes += (long long) (16 * sizeof(double) + 16 * sizeof(int));
  b200md_collect_timers(c);
  memcpy(flags_out, pin_flags, 16 * sizeof(int));
  if (eng_vdwl) *eng_vdwl = eflag ? c->pin_scal.p[0] : 0.0;
  if (virial)
    for (int k = 0; k < 6; k++) virial[k] = vflag ? c->pin_scal.p[1 + k] : 0.0;
  return B200MD_OK;
}

// per-atom energy / virial of one compute call (Pair::eatom, Pair::vatom: accumulated into, like f)
int b200md_peratom_begin(b200md_ctx *c, bool wanted)
{
  c->pa_e = c->pa_v = nullptr;
  if (!wanted) return B200MD_OK;
  const size_t n = (size_t) c->nall;
  CUDA_TRY(c, c->eatom_d.reserve(n + 8));
  CUDA_TRY(c, c->vatom_d.reserve(6 * n + 8));
  CUDA_TRY(c, cudaMemsetAsync(c->eatom_d.p, 0, (n + 8) * sizeof(double), c->stream));
  CUDA_TRY(c, cudaMemsetAsync(c->vatom_d.p, 0, (6 * n + 8) * sizeof(double), c->stream));
  c->pa_e = c->eatom_d.p;
  c->pa_v = c->vatom_d.p;
  return B200MD_OK;
}

int b200md_peratom_finish(b200md_ctx *c, double *eatom, double *vatom)
{
  if (!c->pa_e) return B200MD_OK;
  const size_t n = (size_t) c->nall;
  c->pa_e = c->pa_v = nullptr;
  CUDA_TRY(c, c->pin_pa.reserve(7 * n + 8));
  if (eatom && n) CUDA_TRY(c, cudaMemcpyAsync(c->pin_pa.p, c->eatom_d.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (vatom && n) CUDA_TRY(c, cudaMemcpyAsync(c->pin_pa.p + n, c->vatom_d.p, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (eatom) {
    for (size_t k = 0; k < n; k++) eatom[k] += c->pin_pa.p[k];
    c->d2h_bytes += (long long) (n * sizeof(double));
  }
  if (vatom) {
    for (size_t k = 0; k < 6 * n; k++) vatom[k] += c->pin_pa.p[n + k];
    c->d2h_bytes += (long long) (6 * n * sizeof(double));
  }
  return B200MD_OK;
}

int b200md_collect_timers(b200md_ctx *c)
{
  // caller has synchronised the stream
  for (TimedLaunch &t : c->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
      KernelStat &k = c->kstat[c->kname[t.name_id]];
      k.total_ms += ms;
      k.last_ms = ms;
      k.count++;
    }
    c->event_pool.push_back(t.a);
    c->event_pool.push_back(t.b);
  }
  c->pending.clear();
  return 0;
}

// ---------------------------------------------------------------- atoms
int b200md_upload_atoms(b200md_ctx *c, int nlocal, int nghost, const double *x, const int *type,
                        const int *tag)
{
  ARG_CHECK(c, nlocal >= 0 && nghost >= 0 && ((x && type) || nlocal + nghost == 0), "upload_atoms: bad sizes or NULL arrays");
  int nall = nlocal + nghost;
  ARG_CHECK(c, !c->deterministic || nall <= 32 * (B200MD_DET_BLOCKS - 64),
            "deterministic mode holds per-block partial sums for at most 8.3 M atoms per GPU");
  c->nlocal = nlocal;
  c->nghost = nghost;
  c->nall = nall;
  size_t n = (size_t) nall;
  CUDA_TRY(c, c->x_aos.reserve(3 * n + 8));
  CUDA_TRY(c, c->xq.reserve(n + 8));
  CUDA_TRY(c, c->f.reserve(3 * n + 8));
  CUDA_TRY(c, c->type.reserve(n + 8));
  CUDA_TRY(c, c->tag.reserve(n + 8));
  // types and IDs only change when the host re-sorts or migrates atoms, i.e. together with its neighbor list:
  // they travel once per list hand-over (b200md_set_neighbor_* / b200md_neigh_build reset the marks), positions
  // every call
  if (nall != c->ids_nall) c->type_on_device = c->tag_on_device = false;
  c->ids_nall = nall;
  if (n) {
    CUDA_TRY(c, cudaMemcpyAsync(c->x_aos.p, x, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += (long long) (3 * n * sizeof(double));
    if (!c->type_on_device) {
      CUDA_TRY(c, cudaMemcpyAsync(c->type.p, type, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      c->h2d_bytes += (long long) (n * sizeof(int));
      c->type_on_device = true;
    }
    if (tag && !c->tag_on_device) {
      CUDA_TRY(c, cudaMemcpyAsync(c->tag.p, tag, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      c->h2d_bytes += (long long) (n * sizeof(int));
      c->tag_on_device = true;
    }
  }
  return B200MD_OK;
}

// ---------------------------------------------------------------- exclusive scan (int -> int64, optional alignment)
// Three-phase tile scan; only used when lists are (re)built, never per step.
#define SCAN_BLOCK 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_BLOCK * SCAN_ITEMS)

__device__ __forceinline__ long long scan_elem(const int *in, int i, int n, int align)
{
  if (i >= n) return 0;
  long long v = in[i];
  if (align > 1) v = (v + align - 1) / align * align;
  return v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_sums(const int *__restrict__ in, int n, int align,
                                                            long long *__restrict__ tile_sum)
{
  __shared__ long long sh[SCAN_BLOCK / 32];
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) s += scan_elem(in, base + k, n, align);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < SCAN_BLOCK / 32; w++) t += sh[w];
    tile_sum[blockIdx.x] = t;
  }
}

// one block: thread t owns a contiguous chunk of tiles (a single thread walking 1000 tiles of a 2 M-row list took
// 0.1-0.5 ms of dependent global loads per scan, 1 ms per AEAM rebuild: r02 launch list)
#define SCAN_OFF_BLOCK 1024
__global__ void __launch_bounds__(SCAN_OFF_BLOCK) scan_tile_offsets(long long *tile_sum, int ntiles)
{
  __shared__ long long sh[SCAN_OFF_BLOCK / 32];
  const int per = (ntiles + SCAN_OFF_BLOCK - 1) / SCAN_OFF_BLOCK;
  const int t0 = min(ntiles, (int) threadIdx.x * per), t1 = min(ntiles, t0 + per);
  long long s = 0;
  for (int t = t0; t < t1; t++) s += tile_sum[t];
  long long inc = s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  long long woff = 0;
  for (int w = 0; w < wid; w++) woff += sh[w];
  long long run = woff + inc - s;
  for (int t = t0; t < t1; t++) {
    const long long v = tile_sum[t];
    tile_sum[t] = run;
    run += v;
  }
  if (threadIdx.x == SCAN_OFF_BLOCK - 1) tile_sum[ntiles] = run;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_tile_apply(const int *__restrict__ in, int n, int align,
                                                             const long long *__restrict__ tile_off,
                                                             long long *__restrict__ out)
{
  __shared__ long long sh[SCAN_BLOCK / 32];
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  long long v[SCAN_ITEMS];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = scan_elem(in, base + k, n, align);
    s += v[k];
  }
  // inclusive warp scan of per-thread sums
  long long inc = s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  long long woff = 0;
  for (int w = 0; w < wid; w++) woff += sh[w];
  long long run = tile_off[blockIdx.x] + woff + inc - s;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_off[gridDim.x];
}

int b200md_exclusive_scan_i64(b200md_ctx *c, const int *in, int64_t *out, int n, int align)
{
  if (n <= 0) {
    CUDA_TRY(c, cudaMemsetAsync(out, 0, sizeof(int64_t), c->stream));
    return B200MD_OK;
  }
  int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  CUDA_TRY(c, c->scan_tmp64.reserve((size_t) ntiles + 2));
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  {
    LaunchScope ls(c, "scan");
    scan_tile_sums<<<ntiles, SCAN_BLOCK, 0, c->stream>>>(in, n, align, (long long *) c->scan_tmp64.p);
  }
  {
    LaunchScope ls(c, "scan");
    scan_tile_offsets<<<1, SCAN_OFF_BLOCK, 0, c->stream>>>((long long *) c->scan_tmp64.p, ntiles);
  }
  {
    LaunchScope ls(c, "scan");
    scan_tile_apply<<<ntiles, SCAN_BLOCK, 0, c->stream>>>(in, n, align, (const long long *) c->scan_tmp64.p,
                                                         (long long *) out);
  }
  CUDA_TRY(c, cudaGetLastError());
  return B200MD_OK;
}

// ---------------------------------------------------------------- neighbor list upload
static int finish_list(b200md_ctx *c, int inum, int gnum, int64_t total, double skin)
{
  c->list_inum = inum;
  c->list_gnum = gnum;
  c->list_entries = total;
  c->list_entries_used = total;
  c->skin = skin;
  c->list_valid = true;
  c->inner_valid = false;
  c->type_on_device = c->tag_on_device = false;
  c->n_list_upload++;
  return B200MD_OK;
}

extern "C" int b200md_set_neighbor_list(b200md_ctx *c, int inum, int gnum, const int *numneigh,
                                        const int *const *firstneigh, double skin)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, inum >= 0 && gnum >= 0 && numneigh && firstneigh && skin >= 0.0, "set_neighbor_list");
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rows = inum + gnum;
  std::vector<int64_t> off((size_t) rows + 1);
  int64_t total = 0;
  for (int i = 0; i < rows; i++) {
    off[i] = total;
    total += numneigh[i];
  }
  off[rows] = total;
  CUDA_TRY(c, c->list_off.reserve((size_t) rows + 1));
  CUDA_TRY(c, c->list_num.reserve((size_t) rows + 8));
  CUDA_TRY(c, c->list_val.reserve((size_t) total + 64));
  CUDA_TRY(c, cudaMemcpyAsync(c->list_off.p, off.data(), (rows + 1) * sizeof(int64_t),
                              cudaMemcpyHostToDevice, c->stream));
  if (rows)
    CUDA_TRY(c, cudaMemcpyAsync(c->list_num.p, numneigh, rows * sizeof(int), cudaMemcpyHostToDevice,
                                c->stream));
  // LAMMPS keeps rows in pages: copy every run of rows that is contiguous in host memory at once
  int i = 0;
  while (i < rows) {
    int j = i;
    const int *start = firstneigh[i];
    int64_t len = numneigh[i];
    while (j + 1 < rows && firstneigh[j + 1] == firstneigh[j] + numneigh[j]) {
      j++;
      len += numneigh[j];
    }
    if (len > 0)
      CUDA_TRY(c, cudaMemcpyAsync(c->list_val.p + off[i], start, (size_t) len * sizeof(int),
                                  cudaMemcpyHostToDevice, c->stream));
    i = j + 1;
  }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));    // `off` is a stack-owned staging vector
  c->h2d_bytes += (long long) (total * sizeof(int) + rows * (sizeof(int) + sizeof(int64_t)));
  return finish_list(c, inum, gnum, total, skin);
}

// LAMMPS' NeighList as it really is: numneigh[] and firstneigh[] are indexed by ATOM index, ilist[0..inum+gnum) names
// the atoms that have a row.  The device rows are indexed by atom index as well, so the list must give every owned atom
// (and, with ghost rows, every ghost) exactly one row: skip lists / sub-style lists of pair hybrid are refused.
extern "C" int b200md_set_neighbor_list_ilist(b200md_ctx *c, int inum, int gnum, const int *ilist, const int *numneigh,
                                              const int *const *firstneigh, double skin)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, inum >= 0 && gnum >= 0 && numneigh && firstneigh && skin >= 0.0, "set_neighbor_list_ilist");
  const int rows = inum + gnum;
  if (!ilist) return b200md_set_neighbor_list(c, inum, gnum, numneigh, firstneigh, skin);
  bool identity = true;
  for (int ii = 0; ii < rows && identity; ii++) identity = ilist[ii] == ii;
  if (identity) return b200md_set_neighbor_list(c, inum, gnum, numneigh, firstneigh, skin);
  // a permuted ilist: gather the row descriptors by atom index
  std::vector<int> num((size_t) rows, -1);
  std::vector<const int *> first((size_t) rows, nullptr);
  for (int ii = 0; ii < rows; ii++) {
    const int i = ilist[ii];
    ARG_CHECK(c, i >= 0 && i < rows && num[i] < 0 && (ii < inum) == (i < inum),
              "set_neighbor_list_ilist: the list must hold one row for every owned atom (and every ghost, for ghost "
              "rows): skip lists and sub-style lists are not supported");
    num[i] = numneigh[i];
    first[i] = firstneigh[i];
  }
  return b200md_set_neighbor_list(c, inum, gnum, num.data(), first.data(), skin);
}

extern "C" int b200md_set_neighbor_csr(b200md_ctx *c, int inum, int gnum, const int64_t *offsets,
                                       const int *values, double skin)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, inum >= 0 && gnum >= 0 && offsets && skin >= 0.0, "set_neighbor_csr");
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rows = inum + gnum;
  int64_t total = offsets[rows];
  ARG_CHECK(c, offsets[0] == 0 && total >= 0 && (values || total == 0), "set_neighbor_csr: bad offsets");
  std::vector<int> num((size_t) rows + 1);
  for (int i = 0; i < rows; i++) {
    int64_t n = offsets[i + 1] - offsets[i];
    ARG_CHECK(c, n >= 0 && n < (1 << 30), "set_neighbor_csr: offsets not monotone");
    num[i] = (int) n;
  }
  CUDA_TRY(c, c->list_off.reserve((size_t) rows + 1));
  CUDA_TRY(c, c->list_num.reserve((size_t) rows + 8));
  CUDA_TRY(c, c->list_val.reserve((size_t) total + 64));
  CUDA_TRY(c, cudaMemcpyAsync(c->list_off.p, offsets, (rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                              c->stream));
  if (rows)
    CUDA_TRY(c, cudaMemcpyAsync(c->list_num.p, num.data(), rows * sizeof(int), cudaMemcpyHostToDevice,
                                c->stream));
  if (total)
    CUDA_TRY(c, cudaMemcpyAsync(c->list_val.p, values, (size_t) total * sizeof(int), cudaMemcpyHostToDevice,
                                c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->h2d_bytes += (long long) (total * sizeof(int) + rows * (sizeof(int) + sizeof(int64_t)));
  return finish_list(c, inum, gnum, total, skin);
}

// one plain copy between a host block and the context's device staging, timed with CUDA events on the context's stream
// (measurement: what the link alone needs for the per-step position upload / force download)
extern "C" int b200md_copy_probe(b200md_ctx *c, void *host, size_t bytes, int to_device, double *ms_out)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, host && bytes > 0 && ms_out, "copy_probe: bad arguments");
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, c->x_aos.reserve(bytes / sizeof(double) + 8));
  cudaEvent_t a, b;
  CUDA_TRY(c, cudaEventCreate(&a));
  CUDA_TRY(c, cudaEventCreate(&b));
  CUDA_TRY(c, cudaEventRecord(a, c->stream));
  if (to_device) CUDA_TRY(c, cudaMemcpyAsync(c->x_aos.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
  else CUDA_TRY(c, cudaMemcpyAsync(host, c->x_aos.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaEventRecord(b, c->stream));
  CUDA_TRY(c, cudaEventSynchronize(b));
  float ms = 0.f;
  CUDA_TRY(c, cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *ms_out = ms;
  return B200MD_OK;
}

extern "C" int b200md_neigh_size(b200md_ctx *c, int *nrows, int64_t *nentries)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->list_valid, "neigh_size: no neighbor list on the device");
  if (nrows) *nrows = c->list_inum + c->list_gnum;
  if (nentries) *nentries = c->list_entries;
  return B200MD_OK;
}

extern "C" int b200md_neigh_download(b200md_ctx *c, int *numneigh, int64_t *offsets, int *values)
{
  if (!c) return B200MD_ERR_ARG;
  ARG_CHECK(c, c->list_valid, "neigh_download: no neighbor list on the device");
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rows = c->list_inum + c->list_gnum;
  if (numneigh && rows)
    CUDA_TRY(c, cudaMemcpyAsync(numneigh, c->list_num.p, rows * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (offsets)
    CUDA_TRY(c, cudaMemcpyAsync(offsets, c->list_off.p, (rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                c->stream));
  if (values && c->list_entries)
    CUDA_TRY(c, cudaMemcpyAsync(values, c->list_val.p, (size_t) c->list_entries * sizeof(int),
                                cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200MD_OK;
}
