// twoarmy_b200.cu -- C ABI (include/twoarmy_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/twoarmy_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ta_aux.cuh"
#include "ta_conv1.cuh"
#include "ta_conv1_tc.cuh"
#include "ta_conv1_fwd_ws.cuh"
#include "ta_dgrad_tc.cuh"
#include "ta_feat.cuh"
#include "ta_gae.cuh"
#include "ta_her.cuh"
#include "ta_pred.cuh"
#include "ta_push_tma.cuh"
#include "ta_host.cuh"
#include "ta_stem_bwd_tc.cuh"
#include "ta_step.cuh"
#include "ta_train.cuh"

using namespace ta;

static_assert(sizeof(ta_env_state) == sizeof(EnvStateRec), "public and device record must match");

namespace {

thread_local char g_cuda_err[512] = "";
long long g_launches = 0;

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return TA_E_CUDA;
}
#define CK(call)                                        \
    do {                                                \
        cudaError_t e_ = (call);                        \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

inline int launch_ok(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    g_launches++;
    return TA_OK;
}

inline unsigned blocks_for(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }


}  // namespace

constexpr int HOST_CHUNKS = 8;  // D2H pieces of the packed observations: piece k+1 is on the wire while k is decoded

struct ta_batch {
    int version, view, device, sm_count;
    long long n, npad;
    uint64_t seed, env_id0;
    uint32_t *grid = nullptr;
    uint4 *sc0 = nullptr, *sc1 = nullptr;
    int ctas_per_sm = 0;   // 0 = whatever fits (tuning knob: TA_CTAS_PER_SM)
    int pdl = 3;           // programmatic dependent launch of the step kernel (TA_PDL): 0 off; 1 trigger at the top of the
                           // kernel; 2 trigger after the kernel's own dependency wait + early state load when the previous
                           // writer on the stream was another handle; 3 (default) the same with the trigger after the last
                           // observation pass.  Measured on B200, 65536 envs, graph-replayed single-step launches over 8
                           // rotating batches: 16.0 / 15.2 / 14.4 / 14.3 us (V = 17), 8.85 / 8.02 / 7.27 / 7.26 us (V = 7)
    int warps_per_cta = 0; // independent tiles per CTA; 0 = chosen per launch (TA_WARPS_PER_CTA, 1..8)
    int debug_flags = 0;   // TA_DEBUG_FLAGS: timing experiments (StepArgs::flags bits 2,3)
    uint32_t *tmpl = nullptr;  // [20] the _gen_grid record
    // host-call path (ta_step_host)
    cudaStream_t own_stream = nullptr;
    void *d_act = nullptr;
    uint8_t *d_obs = nullptr;        // expanded observations (TA_STEP_HOST_DMA)
    uint32_t *d_codes = nullptr;     // packed observations [ntiles][2*V*V] followed by the status bytes [npad]
    uint8_t *h_codes = nullptr;      // pinned staging of the same
    cudaEvent_t chunk_ev[HOST_CHUNKS] = {};
    ta_host::Pool *pool = nullptr;
    ta_host::Job job;
    long long last_d2h_bytes = 0;
    int auto_dma = -1;               // -1 undecided, 0 host threads expand the packed form, 1 DMA of the expanded form
    // timing
    int timing = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

// Which handle's state did the kernel this library launched last on a stream write?  The step kernel may load its
// first tile's state ahead of the programmatic-dependency wait only if that was a different handle (ta_step.cuh).
struct StreamWriter { cudaStream_t stream; const ta_batch *handle; };
StreamWriter g_writers[32];
int g_nwriters = 0;
bool g_writers_overflow = false;
const ta_batch *last_writer(cudaStream_t st) {
    for (int i = 0; i < g_nwriters; i++)
        if (g_writers[i].stream == st) return g_writers[i].handle;
    return nullptr;
}
void note_writer(cudaStream_t st, const ta_batch *h) {
    for (int i = 0; i < g_nwriters; i++)
        if (g_writers[i].stream == st) { g_writers[i].handle = h; return; }
    if (g_nwriters < 32) g_writers[g_nwriters++] = {st, h};
    // table full: a stream we forget could still have this handle's writer in flight, so stop loading early for good
    else g_writers_overflow = true;
}

// the _gen_grid record (twoarmy_v4.py:38-80); padding cells 289..319 hold the wall code
void build_template(uint32_t *tm) {
    for (int w = 0; w < REC_WORDS; w++) tm[w] = 0x55555555u;
    for (int x = 0; x < GS; x++)
        for (int y = 0; y < GS; y++) cell_set(tm, x, y, initial_cell(x, y));
}

template <int V>
int launch_step_t(ta_batch *h, const StepArgs &a_in, cudaStream_t st) {
    StepArgs a = a_in;
    if (h->pdl >= 2) {  // trigger after the wait (2) / after the last observation pass (3); early state load when safe
        a.flags |= (h->pdl == 2 ? 1 : 2) << 7;
        const ta_batch *prev = last_writer(st);
        if (!g_writers_overflow && prev != nullptr && prev != h) a.flags |= 64;
    }
    if (!(a.flags & 16)) note_writer(st, h);
    static int ctas_per_sm[64][STEP_MAX_WARPS + 1] = {};
    auto kern = step_obs_kernel<V>;
    // warps (= independent 32-env tiles) per CTA: whatever spreads the tiles most evenly over the
    // SMs, the larger count on a tie (fewer CTAs to dispatch; measured equal or slightly better)
    int W = h->warps_per_cta;
    if (W <= 0) {
        long long best = -1;
        for (int w = 1; w <= STEP_MAX_WARPS; w++) {
            const long long ctas = (a.ntiles + w - 1) / w;
            const long long worst = (ctas + h->sm_count - 1) / h->sm_count * w;  // tiles on the fullest SM
            if (best < 0 || worst <= best) { best = worst; W = w; }
        }
    }
    const int smem = ObsCfg<V>::SMEM * W;
    int &cps = ctas_per_sm[h->device & 63][W];
    if (!cps) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ObsCfg<V>::SMEM * STEP_MAX_WARPS));
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * W, smem));
        cps = occ > 0 ? occ : 1;
    }
    int per_sm = cps;
    if (h->ctas_per_sm > 0 && h->ctas_per_sm < per_sm) per_sm = h->ctas_per_sm;
    const int max_ctas = h->sm_count * per_sm;
    const int want = (a.ntiles + W - 1) / W;
    int grid = want < max_ctas ? want : max_ctas;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(32 * W);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;  // (without the attribute the kernel's griddepcontrol instructions are no-ops)
    CK(cudaLaunchKernelEx(&cfg, kern, a));
    return launch_ok("step_obs_kernel");
}

int g_force_generic = 0;

int launch_step(ta_batch *h, const StepArgs &a, cudaStream_t st) {
    switch (h->view) {
        case 3: return launch_step_t<3>(h, a, st);
        case 5: return launch_step_t<5>(h, a, st);
        case 7: return launch_step_t<7>(h, a, st);
        case 9: return launch_step_t<9>(h, a, st);
        case 11: return launch_step_t<11>(h, a, st);
        case 13: return launch_step_t<13>(h, a, st);
        case 15: return launch_step_t<15>(h, a, st);
        case 17: return launch_step_t<17>(h, a, st);
    }
    return TA_E_UNSUPPORTED;
}

// gen_obs() of the current state: the step kernel's obs pass alone (flags bit 4)
int launch_observe(ta_batch *h, uint8_t *obs_out, cudaStream_t st) {
    if ((uintptr_t)obs_out & 15u) {  // unaligned caller buffer: the simple per-cell kernel
        observe_kernel<<<blocks_for(h->n * h->view * h->view, 256), 256, 0, st>>>(h->grid, h->sc0, obs_out, h->view, h->n);
        return launch_ok("observe_kernel");
    }
    StepArgs a = {};
    a.grid = h->grid; a.sc0 = h->sc0; a.sc1 = h->sc1; a.tmpl = h->tmpl;
    a.obs = obs_out;
    a.n = h->n; a.ntiles = (int)(h->npad / TILE); a.T = 1;
    a.version = h->version; a.flags = 16 | (g_force_generic ? 2 : 0);
    a.env_id0 = h->env_id0;
    return launch_step(h, a, st);
}

int do_reset(ta_batch *h, const uint8_t *mask, int hard, uint8_t *obs_out, cudaStream_t st, bool pad_too) {
    // padded tail envs (>= n) are reset only at creation
    const long long cnt = pad_too ? h->npad : h->n;
    note_writer(st, h);
    reset_grid_kernel<<<blocks_for(cnt * REC_WORDS, 256), 256, 0, st>>>(
        h->grid, h->tmpl, mask, cnt);
    if (int rc = launch_ok("reset_grid_kernel")) return rc;
    reset_scalar_kernel<<<blocks_for(cnt, 256), 256, 0, st>>>(h->sc0, h->sc1, mask, hard, cnt);
    if (int rc = launch_ok("reset_scalar_kernel")) return rc;
    if (obs_out) return launch_observe(h, obs_out, st);
    return TA_OK;
}

}  // namespace

extern "C" {

int ta_abi_version(void) { return TA_ABI_VERSION; }

const char *ta_strerror(int code) {
    switch (code) {
        case TA_OK: return "ok";
        case TA_E_INVALID: return "invalid argument";
        case TA_E_CUDA: return "CUDA runtime error (see ta_last_cuda_error)";
        case TA_E_NOMEM: return "out of memory";
        case TA_E_UNSUPPORTED: return "unsupported configuration";
    }
    return "unknown error";
}

const char *ta_last_cuda_error(void) { return g_cuda_err; }
int64_t ta_launch_count(void) { return g_launches; }

int ta_create(ta_handle *out, int version, int64_t n_envs, int view, int device, uint64_t seed, uint64_t env_id0) {
    if (!out || (version != 4 && version != 6) || n_envs <= 0 || view < 3 || view > 17 || (view & 1) == 0)
        return TA_E_INVALID;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return TA_E_INVALID;
    CK(cudaSetDevice(device));
    ta_batch *h = new (std::nothrow) ta_batch();
    if (!h) return TA_E_NOMEM;
    h->version = version; h->view = view; h->device = device;
    h->n = n_envs; h->npad = (n_envs + TILE - 1) / TILE * TILE;
    h->seed = seed; h->env_id0 = env_id0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("TA_WARPS_PER_CTA")) {
        int v = atoi(e);
        if (v >= 1 && v <= STEP_MAX_WARPS) h->warps_per_cta = v;
    }
    if (const char *e = getenv("TA_PDL")) h->pdl = atoi(e) >= 0 && atoi(e) <= 3 ? atoi(e) : 0;
    if (const char *e = getenv("TA_DEBUG_FLAGS")) h->debug_flags = atoi(e) & 3;
    if (const char *e = getenv("TA_CTAS_PER_SM")) {  // tuning knob for experiments
        int v = atoi(e);
        if (v >= 1 && v <= 32) h->ctas_per_sm = v;
    }
    if (prop.major < 10) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
        delete h;
        return TA_E_UNSUPPORTED;
    }
    CK(cudaMalloc(&h->grid, (size_t)h->npad * REC_BYTES));
    CK(cudaMalloc(&h->sc0, (size_t)h->npad * sizeof(uint4)));
    CK(cudaMalloc(&h->sc1, (size_t)h->npad * sizeof(uint4)));
    CK(cudaMalloc(&h->tmpl, REC_BYTES + 16));
    CK(cudaMemset(h->sc0, 0, (size_t)h->npad * sizeof(uint4)));
    CK(cudaMemset(h->sc1, 0, (size_t)h->npad * sizeof(uint4)));
    uint32_t tm[REC_WORDS + 4];
    build_template(tm);
    tm[REC_WORDS] = tm[REC_WORDS + 2] = TYPE_LUT;
    tm[REC_WORDS + 1] = tm[REC_WORDS + 3] = COLOR_LUT;
    CK(cudaMemcpy(h->tmpl, tm, REC_BYTES + 16, cudaMemcpyHostToDevice));
    CK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    if (int rc = do_reset(h, nullptr, 1, nullptr, nullptr, true)) return rc;
    CK(cudaDeviceSynchronize());
    *out = h;
    return TA_OK;
}

int ta_destroy(ta_handle h) {
    if (!h) return TA_E_INVALID;
    cudaSetDevice(h->device);
    cudaFree(h->grid); cudaFree(h->sc0); cudaFree(h->sc1); cudaFree(h->tmpl);
    cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_codes);
    if (h->h_codes) cudaFreeHost(h->h_codes);
    for (int c = 0; c < HOST_CHUNKS; c++)
        if (h->chunk_ev[c]) cudaEventDestroy(h->chunk_ev[c]);
    delete h->pool;
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return TA_OK;
}

int64_t ta_num_envs(ta_handle h) { return h ? h->n : 0; }
int ta_view(ta_handle h) { return h ? h->view : 0; }
int ta_version(ta_handle h) { return h ? h->version : 0; }

int ta_reset(ta_handle h, const uint8_t *mask, int hard, uint8_t *obs_out, void *stream) {
    if (!h) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    return do_reset(h, mask, hard, obs_out, (cudaStream_t)stream, false);
}

int ta_observe_general(ta_handle h, const uint8_t *agent_dirs, int agent_dir, const uint8_t *see_through, int see_through_all,
                       uint8_t *obs_out, void *stream) {
    if (!h || !obs_out || (!agent_dirs && (agent_dir < 0 || agent_dir > 3))) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    observe_general_kernel<<<blocks_for(h->n, 128), 128, 0, (cudaStream_t)stream>>>(h->grid, h->sc0, agent_dirs, agent_dir, see_through,
                                                                                   see_through_all, obs_out, h->view, h->n);
    return launch_ok("observe_general_kernel");
}

static int step_launch(ta_handle h, const void *actions, int action_dtype, const uint8_t *draws, int flags, int T,
                       uint8_t *obs_out, float *reward_out, uint8_t *term_out, uint8_t *trunc_out, uint8_t *consumed_out,
                       void *stream, uint8_t *status_out = nullptr, bool packed_obs = false) {
    if (!h || !actions || !obs_out || T <= 0) return TA_E_INVALID;
    if (!status_out && (!reward_out || !term_out || !trunc_out)) return TA_E_INVALID;
    if (action_dtype < 0 || action_dtype > 2) return TA_E_INVALID;
    if (((uintptr_t)obs_out & 15u) || (draws && ((uintptr_t)draws & 7u))) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    StepArgs a = {};
    a.grid = h->grid; a.sc0 = h->sc0; a.sc1 = h->sc1; a.tmpl = h->tmpl;
    a.actions = actions; a.draws = draws;
    a.obs = obs_out; a.reward = reward_out; a.term = term_out; a.trunc = trunc_out; a.consumed = consumed_out;
    a.status = status_out;
    a.n = h->n; a.ntiles = (int)(h->npad / TILE); a.T = T;
    a.version = h->version;
    a.flags = (flags & 1) | (g_force_generic ? 2 : 0) | (h->debug_flags << 2) | (packed_obs ? 32 : 0);
    a.action_dtype = action_dtype;
    a.seed_lo = (uint32_t)h->seed; a.seed_hi = (uint32_t)(h->seed >> 32);
    a.env_id0 = h->env_id0;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->timing) CK(cudaEventRecord(h->ev0, st));
    int rc = launch_step(h, a, st);
    if (h->timing) CK(cudaEventRecord(h->ev1, st));
    return rc;
}

int ta_step(ta_handle h, const void *actions, int action_dtype, const uint8_t *draws, int flags, uint8_t *obs_out,
            float *reward_out, uint8_t *term_out, uint8_t *trunc_out, uint8_t *consumed_out, void *stream) {
    return step_launch(h, actions, action_dtype, draws, flags, 1, obs_out, reward_out, term_out, trunc_out, consumed_out,
                       stream);
}

int ta_step_packed(ta_handle h, const void *actions, int action_dtype, const uint8_t *draws, int flags, uint32_t *codes_out,
                   uint8_t *status_out, uint8_t *consumed_out, void *stream) {
    if (!status_out || ((uintptr_t)codes_out & 3u)) return TA_E_INVALID;
    return step_launch(h, actions, action_dtype, draws, flags, 1, (uint8_t *)codes_out, nullptr, nullptr, nullptr, consumed_out, stream,
                       status_out, true);
}

// ta_step_host: H2D actions -> fused kernel -> D2H results -> (packed form) decode on the host threads
int ta_step_host(ta_handle h, const void *actions, int action_dtype, int flags, uint8_t *obs_out, float *reward_out,
                 uint8_t *term_out, uint8_t *trunc_out) {
    if (!h || !actions || !obs_out || !reward_out || !term_out || !trunc_out) return TA_E_INVALID;
    if (action_dtype < 0 || action_dtype > 2) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    const size_t asz = action_dtype == TA_ACT_I32 ? 4 : (action_dtype == TA_ACT_U8 ? 1 : 8);
    const int V = h->view, runs = 2 * V * V, obs_bytes = 3 * V * V;
    const long long ntiles = h->npad / TILE;
    const size_t codes_bytes = (size_t)ntiles * runs * 4, status_bytes = (size_t)h->npad;
    if (!h->d_codes) {
        CK(cudaMalloc(&h->d_act, (size_t)h->n * 8));
        CK(cudaMalloc(&h->d_codes, codes_bytes + status_bytes));
        CK(cudaHostAlloc(&h->h_codes, codes_bytes + status_bytes, cudaHostAllocDefault));
        for (int c = 0; c < HOST_CHUNKS; c++) CK(cudaEventCreateWithFlags(&h->chunk_ev[c], cudaEventDisableTiming));
    }
    cudaStream_t st = h->own_stream;
    uint8_t *d_status = (uint8_t *)h->d_codes + codes_bytes, *h_status = h->h_codes + codes_bytes;
    CK(cudaMemcpyAsync(h->d_act, actions, (size_t)h->n * asz, cudaMemcpyHostToDevice, st));
    ta_host::Job &job = h->job;
    job.status = h_status; job.reward = reward_out; job.term = term_out; job.trunc = trunc_out;
    job.n = h->n; job.ntiles = ntiles; job.runs = runs; job.obs_bytes = obs_bytes; job.obs = obs_out;
    // A single rank on a host with few cores and a pinned destination: the DMA engine alone (~54 GB/s of expanded
    // observations) beats the threads' expansion (~11 GB/s each).  With several ranks on the node the DMA form does not
    // scale (measured on 8 GPUs / 32 cores: 111 M env-steps/s for all ranks together against 196 M through the packed
    // form with 4 threads per rank), so it is never chosen there.  TA_STEP_HOST_AUTO=0 disables the choice.
    if (!(flags & TA_STEP_HOST_DMA) && h->auto_dma < 0) {
        const char *e = getenv("TA_STEP_HOST_AUTO"), *lw = getenv("LOCAL_WORLD_SIZE");
        h->auto_dma = 0;
        if (!(e && atoi(e) == 0) && !(lw && atoi(lw) > 1) && ta_host::default_threads() <= 4) {
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, obs_out) == cudaSuccess && pa.type == cudaMemoryTypeHost) h->auto_dma = 1;
            (void)cudaGetLastError();
        }
    }
    if (h->auto_dma > 0) flags |= TA_STEP_HOST_DMA;
    if (flags & TA_STEP_HOST_DMA) {
        // the expanded observations straight over PCIe into the caller's array (pinned: one DMA); the status bytes
        // (reward index, terminated, truncated) through the staging buffer, decoded here
        if (!h->d_obs) CK(cudaMalloc(&h->d_obs, (size_t)h->n * obs_bytes));
        if (int rc = step_launch(h, h->d_act, action_dtype, nullptr, flags & 1, 1, h->d_obs, nullptr, nullptr, nullptr, nullptr, st,
                                 d_status, false))
            return rc;
        CK(cudaMemcpyAsync(obs_out, h->d_obs, (size_t)h->n * obs_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_status, d_status, status_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (long long e = 0; e < h->n; e++) {
            const uint8_t s = h_status[e];
            memcpy(reward_out + e, &ta_host::REWARD_BITS[s & 7u], 4);
            term_out[e] = (s >> 3) & 1u;
            trunc_out[e] = (s >> 4) & 1u;
        }
        h->last_d2h_bytes = (long long)h->n * obs_bytes + (long long)status_bytes;
        return TA_OK;
    }
    if (!h->pool) {
        const int nt = ta_host::default_threads();
        h->pool = new (std::nothrow) ta_host::Pool(nt > 1 ? nt - 1 : 0);  // the calling thread decodes too
        if (!h->pool) return TA_E_NOMEM;
    }
    if (int rc = step_launch(h, h->d_act, action_dtype, nullptr, flags & 1, 1, (uint8_t *)h->d_codes, nullptr, nullptr, nullptr, nullptr, st,
                             d_status, true))
        return rc;
    // status first (small), then the packed observations in HOST_CHUNKS pieces, an event behind each
    CK(cudaMemcpyAsync(h_status, d_status, status_bytes, cudaMemcpyDeviceToHost, st));
    long long unit_tiles = 32768 / (TILE * obs_bytes);
    if (unit_tiles < 1) unit_tiles = 1;
    job.unit_tiles = (int)unit_tiles;
    job.nunits = (ntiles + unit_tiles - 1) / unit_tiles;
    job.codes = reinterpret_cast<const uint32_t *>(h->h_codes);
    job.next.store(0); job.ready.store(0); job.finished.store(0);
    long long chunk_end[HOST_CHUNKS];
    int nchunks = 0;
    for (int c = 0; c < HOST_CHUNKS; c++) {
        long long u0 = job.nunits * c / HOST_CHUNKS, u1 = job.nunits * (c + 1) / HOST_CHUNKS;
        if (u1 == u0) continue;
        long long t0 = u0 * unit_tiles, t1 = u1 * unit_tiles < ntiles ? u1 * unit_tiles : ntiles;
        CK(cudaMemcpyAsync(h->h_codes + (size_t)t0 * runs * 4, (const uint8_t *)h->d_codes + (size_t)t0 * runs * 4, (size_t)(t1 - t0) * runs * 4,
                           cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(h->chunk_ev[nchunks], st));
        chunk_end[nchunks++] = u1;
    }
    h->pool->start(&job);
    cudaError_t err = cudaSuccess;
    for (int c = 0; c < nchunks; c++) {
        const cudaError_t e = cudaEventSynchronize(h->chunk_ev[c]);
        if (e != cudaSuccess) err = e;
        job.ready.store(chunk_end[c], std::memory_order_release);  // (on an error too: the workers must not spin forever)
    }
    h->pool->finish(&job);
    if (err != cudaSuccess) return cuda_fail(err, "ta_step_host: D2H");
    h->last_d2h_bytes = (long long)(codes_bytes + status_bytes);
    return TA_OK;
}

int64_t ta_step_host_d2h_bytes(ta_handle h) { return h ? h->last_d2h_bytes : 0; }

int ta_decode_packed_host(const uint32_t *codes, const uint8_t *status, int64_t n, int view, uint8_t *obs_out, float *reward_out,
                          uint8_t *term_out, uint8_t *trunc_out) {
    if (!codes || !status || !obs_out || !reward_out || !term_out || !trunc_out || n <= 0 || view < 3 || view > 17 || !(view & 1))
        return TA_E_INVALID;
    ta_host::Job job;
    job.codes = codes; job.status = status; job.obs = obs_out; job.reward = reward_out; job.term = term_out; job.trunc = trunc_out;
    job.n = n; job.ntiles = (n + TILE - 1) / TILE; job.runs = 2 * view * view; job.obs_bytes = 3 * view * view;
    job.unit_tiles = 1; job.nunits = job.ntiles;
    job.ready.store(job.nunits);
    ta_host::work(job);
    return TA_OK;
}

int ta_rollout(ta_handle h, const void *actions, int action_dtype, int T, uint8_t *obs_out, float *reward_out,
               uint8_t *term_out, uint8_t *trunc_out, void *stream) {
    // one persistent launch: the env state stays on chip for all T steps
    return step_launch(h, actions, action_dtype, nullptr, TA_STEP_AUTORESET, T, obs_out, reward_out, term_out, trunc_out,
                       nullptr, stream);
}

}  // extern "C"
namespace {
int g_push_tma = -1;   // -1: read TA_PUSH_TMA on first use (default 1); ta_debug_push_tma() sets it
int tc_fail_flag(int **out);

// the persistent TMA form of the uint8 frame-stack push (ta_push_tma.cuh); *used = 0 when its preconditions do not hold
int launch_push_tma(cudaStream_t st, const uint32_t *grid, const uint4 *sc0, const uint8_t *s_prev, uint8_t *s_out, const float *p_prev,
                    float *p_out, const uint8_t *prev_done, long long n, int *used) {
    *used = 0;
    static int stages = 0, ctas = 0;   // TA_PUSH_TMA (default 1), TA_PUSH_STAGES (2..4), TA_PUSH_CTAS (per SM; 0 = what fits)
    int &mode = g_push_tma;
    if (mode < 0) { const char *e = getenv("TA_PUSH_TMA"); mode = e ? atoi(e) : 1; }
    if (!stages) {
        const char *e;
        stages = 3;
        if ((e = getenv("TA_PUSH_STAGES"))) stages = atoi(e);
        if (stages < 2 || stages > 4) stages = 3;
        if ((e = getenv("TA_PUSH_CTAS"))) ctas = atoi(e);
    }
    if (!mode || n < FEAT_ENVS || (n % FEAT_ENVS) != 0) return TA_OK;
    if ((((uintptr_t)s_prev | (uintptr_t)s_out | (uintptr_t)prev_done | (uintptr_t)grid) & 15u) != 0) return TA_OK;
    int *fail = nullptr;
    if (int rc = tc_fail_flag(&fail)) return rc;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int smem = stages * PT_STAGE;
    const void *kern = stages == 2 ? (const void *)stack_push_tma_kernel<2> : stages == 3 ? (const void *)stack_push_tma_kernel<3>
                                                                                       : (const void *)stack_push_tma_kernel<4>;
    static bool attr_set[64][5] = {};
    if (!attr_set[dev & 63][stages]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set[dev & 63][stages] = true;
    }
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PT_THREADS, smem));
    if (occ < 1) return TA_OK;
    // two resident CTAs per SM by default: measured 34.0 us per 65536 envs against 35.5 with three (4096 tiles over 296 CTAs
    // are 13-14 each, over 444 CTAs 9-10: the static round-robin's tail costs more than the third CTA hides)
    const int want = ctas > 0 ? ctas : 2;
    if (want < occ) occ = want;
    const long long ntiles = n / FEAT_ENVS;
    long long g = (long long)sms * occ;
    if (g > ntiles) g = ntiles;
    if (stages == 2) stack_push_tma_kernel<2><<<(unsigned)g, PT_THREADS, smem, st>>>(grid, sc0, s_prev, s_out, p_prev, p_out, prev_done, n, fail);
    else if (stages == 3) stack_push_tma_kernel<3><<<(unsigned)g, PT_THREADS, smem, st>>>(grid, sc0, s_prev, s_out, p_prev, p_out, prev_done, n, fail);
    else stack_push_tma_kernel<4><<<(unsigned)g, PT_THREADS, smem, st>>>(grid, sc0, s_prev, s_out, p_prev, p_out, prev_done, n, fail);
    *used = 1;
    return TA_OK;
}
}  // namespace
extern "C" {

// the uint8 frame-stack push; TA_PUSH_MINB (tuning knob) = CTAs per SM the kernel's registers are bounded for
static void launch_push_codes(unsigned nb, cudaStream_t st, const uint32_t *grid, const uint4 *sc0, const uint8_t *s_prev,
                              uint8_t *s_out, const float *p_prev, float *p_out, const uint8_t *prev_done, int init_all,
                              long long n) {
    static int minb = -1;
    if (minb < 0) { const char *e = getenv("TA_PUSH_MINB"); minb = e ? atoi(e) : 5; }
    static int dbg = -1;   // TA_PUSH_DBG: ablation switches of scripts/probe_push.py (results undefined when set)
    if (dbg < 0) { const char *e = getenv("TA_PUSH_DBG"); dbg = e ? atoi(e) : 0; }
    init_all = (init_all ? 1 : 0) | (dbg << 8);
    static bool carved[64] = {};   // per device: ask for the largest shared-memory carveout (8 CTAs x 25 KB per SM)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!carved[dev & 63]) {
        cudaFuncSetAttribute(stack_push_codes_tile_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(stack_push_codes_tile_kernel<6>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        cudaFuncSetAttribute(stack_push_codes_tile_kernel<5>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        carved[dev & 63] = true;
    }
    static int extra = -1;  // TA_PUSH_EXTRA_SMEM: unused dynamic shared memory per CTA (occupancy throttle of the probe)
    if (extra < 0) { const char *e = getenv("TA_PUSH_EXTRA_SMEM"); extra = e ? atoi(e) : 0; if (extra < 0 || extra > 22000) extra = 0; }
    if (minb >= 8) stack_push_codes_tile_kernel<8><<<nb, FEAT_THREADS, extra, st>>>(grid, sc0, s_prev, s_out, p_prev, p_out, prev_done, init_all, n);
    else if (minb >= 6) stack_push_codes_tile_kernel<6><<<nb, FEAT_THREADS, extra, st>>>(grid, sc0, s_prev, s_out, p_prev, p_out, prev_done, init_all, n);
    else stack_push_codes_tile_kernel<5><<<nb, FEAT_THREADS, extra, st>>>(grid, sc0, s_prev, s_out, p_prev, p_out, prev_done, init_all, n);
}

int ta_state_matrix(ta_handle h, uint8_t *codes_out, float *matrix_out, float *place_out, void *stream) {
    if (!h) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    if ((((uintptr_t)codes_out | (uintptr_t)matrix_out) & 15u) == 0 && h->n >= FEAT_ENVS && (h->n % FEAT_ENVS) == 0) {
        // the pipelined form (ta_push_tma.cuh); TA_FEAT_PIPE=0 / ta_debug_push_tma(0) select the one-tile-per-CTA kernel
        int &mode = g_push_tma;
        if (mode < 0) { const char *e = getenv("TA_PUSH_TMA"); mode = e ? atoi(e) : 1; }
        static int pipe = -1, ctas = 6;   // (measured: 12.5 us with 4 CTAs per SM, 11.7 with 6 or 7, per 65536 envs)
        if (pipe < 0) {
            const char *e = getenv("TA_FEAT_PIPE"); pipe = e ? atoi(e) : 1;
            if ((e = getenv("TA_FEAT_CTAS")) && atoi(e) >= 1 && atoi(e) <= 7) ctas = atoi(e);
        }
        if (mode && pipe) {
            int *fail = nullptr;
            if (int rc = tc_fail_flag(&fail)) return rc;
            int dev = 0, sms = 0;
            CK(cudaGetDevice(&dev));
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            long long g = (long long)sms * ctas;
            if (g > h->n / FEAT_ENVS) g = h->n / FEAT_ENVS;
            frame_codes_pipe_kernel<<<(unsigned)g, PT_THREADS, 0, (cudaStream_t)stream>>>(h->grid, h->sc0, codes_out, matrix_out, place_out,
                                                                                          h->n, fail);
            return launch_ok("frame_codes_pipe_kernel");
        }
    }
    if ((((uintptr_t)codes_out | (uintptr_t)matrix_out) & 15u) == 0) {
        frame_codes_tile_kernel<<<blocks_for(h->n, FEAT_ENVS), FEAT_THREADS, 0, (cudaStream_t)stream>>>(
            h->grid, h->sc0, codes_out, matrix_out, place_out, h->n);
        return launch_ok("frame_codes_tile_kernel");
    }
    state_matrix_kernel<<<blocks_for(h->n, SM_ENVS), 320, 0, (cudaStream_t)stream>>>(h->grid, h->sc0, codes_out, matrix_out,
                                                                                   place_out, h->n);
    return launch_ok("state_matrix_kernel");
}

int ta_stack_roll(ta_handle h, float *s_stack, float *p_stack, const uint8_t *init_mask, int init, void *stream) {
    if (!h || !s_stack) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    stack_roll_kernel<float><<<blocks_for(h->n, SM_ENVS), 320, 0, (cudaStream_t)stream>>>(h->grid, h->sc0, s_stack, p_stack,
                                                                                        init_mask, init, h->n);
    return launch_ok("stack_roll_kernel");
}

int ta_stack_roll_codes(ta_handle h, uint8_t *s_codes, float *p_stack, const uint8_t *init_mask, int init, void *stream) {
    if (!h || !s_codes) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    if (!init && ((uintptr_t)s_codes & 15u) == 0) {  // the roll itself: the tile kernel, in place
        int used = 0;
        if (int rc = launch_push_tma((cudaStream_t)stream, h->grid, h->sc0, s_codes, s_codes, p_stack, p_stack, nullptr, h->n, &used)) return rc;
        if (used) return launch_ok("stack_push_tma_kernel");
        launch_push_codes(blocks_for(h->n, FEAT_ENVS), (cudaStream_t)stream, h->grid, h->sc0, s_codes, s_codes, p_stack, p_stack,
                          nullptr, 0, h->n);
        return launch_ok("stack_push_codes_tile_kernel");
    }
    stack_roll_kernel<uint8_t><<<blocks_for(h->n, SM_ENVS), 320, 0, (cudaStream_t)stream>>>(h->grid, h->sc0, s_codes, p_stack,
                                                                                          init_mask, init, h->n);
    return launch_ok("stack_roll_kernel<u8>");
}

int ta_stack_push(ta_handle h, const void *s_prev, void *s_out, const float *p_prev, float *p_out, const uint8_t *prev_done,
                  int init_all, int dtype, void *stream) {
    if (!h || !s_out || (dtype != TA_STACK_F32 && dtype != TA_STACK_U8)) return TA_E_INVALID;
    if (!init_all && (!s_prev || (p_out && !p_prev))) return TA_E_INVALID;
    if (s_prev == s_out || (p_out && p_prev == p_out)) return TA_E_INVALID;  // out of place only
    if ((((uintptr_t)s_prev | (uintptr_t)s_out) & 15u) != 0) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    const unsigned nb = blocks_for(h->n, FEAT_ENVS);
    if (dtype == TA_STACK_U8 && !init_all) {
        int used = 0;
        if (int rc = launch_push_tma((cudaStream_t)stream, h->grid, h->sc0, (const uint8_t *)s_prev, (uint8_t *)s_out, p_prev, p_out,
                                     prev_done, h->n, &used)) return rc;
        if (used) return launch_ok("stack_push_tma_kernel");
    }
    if (dtype == TA_STACK_U8)
        launch_push_codes(nb, (cudaStream_t)stream, h->grid, h->sc0, (const uint8_t *)s_prev, (uint8_t *)s_out, p_prev, p_out,
                          prev_done, init_all, h->n);
    else
        stack_push_kernel<float><<<nb, FEAT_THREADS, 0, (cudaStream_t)stream>>>(
            h->grid, h->sc0, (const float *)s_prev, (float *)s_out, p_prev, p_out, prev_done, init_all, h->n);
    return launch_ok("stack_push_kernel");
}

int ta_render(ta_handle h, const uint8_t *atlas, int tile_size, int highlight, const int64_t *env_ids, int64_t m,
              uint8_t *rgb_out, void *stream) {
    if (!h || !atlas || !rgb_out || tile_size < 1 || tile_size > 64 || m <= 0) return TA_E_INVALID;
    if (!env_ids && m > h->n) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    const long long total = (long long)m * GS * tile_size * GS * tile_size * 3;
    unsigned nb = blocks_for(total, 256 * 4);
    if (nb > 148u * 32u) nb = 148u * 32u;
    render_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(h->grid, h->sc0, atlas, (const long long *)env_ids, m, tile_size,
                                                      highlight, h->view, rgb_out);
    return launch_ok("render_kernel");
}

int ta_export_state(ta_handle h, ta_env_state *out, void *stream) {
    if (!h || !out) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    export_kernel<<<blocks_for(h->n * NCELL, 256), 256, 0, (cudaStream_t)stream>>>(h->grid, h->sc0, h->sc1,
                                                                                 reinterpret_cast<EnvStateRec *>(out), h->n);
    return launch_ok("export_kernel");
}

int ta_import_state(ta_handle h, const ta_env_state *in, void *stream) {
    if (!h || !in) return TA_E_INVALID;
    CK(cudaSetDevice(h->device));
    note_writer((cudaStream_t)stream, h);
    import_kernel<<<blocks_for(h->n * REC_WORDS, 256), 256, 0, (cudaStream_t)stream>>>(
        h->grid, h->sc0, h->sc1, reinterpret_cast<const EnvStateRec *>(in), h->n);
    return launch_ok("import_kernel");
}

// coop (nullable): a zeroed ticket word -- the launch then also normalises adv (single launch; only taken when one pass
// covers T and the whole grid is resident at once, *coop_used says whether it was)
static int gae_launch(const float *reward, const float *v, const float *v_next, const float *last_v, const uint8_t *done, float gamma,
                      float lam, int use_mask, int T, int64_t n, float *adv_out, float *ret_out, double *stats3, void *stream,
                      unsigned int *coop = nullptr, int *coop_used = nullptr) {
    if (coop_used) *coop_used = 0;
    if (!reward || !v || !adv_out || !ret_out || T <= 0 || n <= 0) return TA_E_INVALID;
    if (!v_next && !last_v) return TA_E_INVALID;
    if (use_mask && !done) return TA_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (stats3) CK(cudaMemsetAsync(stats3, 0, (coop ? 5 : 3) * sizeof(double), st));   // (coop = stats3 + 3: the ticket word)
    const uintptr_t al = (uintptr_t)reward | (uintptr_t)v | (uintptr_t)v_next | (uintptr_t)last_v | (uintptr_t)adv_out |
                         (uintptr_t)ret_out;
    if ((n & 3) == 0 && (al & 15u) == 0 && ((uintptr_t)done & 3u) == 0) {  // four envs per thread, 16-byte accesses
        // 8 steps per thread; time chunks per CTA: TA_GAE_CH (tuning knob)
        // measured on B200 (T = 128): 4 chunks (32 steps a pass, 128-thread CTAs) reach 0.89 of the HBM
        // peak once there are enough CTAs; small problems want the extra time-parallelism of 8
        static int gch = -1;
        if (gch < 0) { const char *e = getenv("TA_GAE_CH"); gch = e ? atoi(e) : 0; if (gch < 0 || gch > 16) gch = 0; }
        const int steps_chunks = (T + 7) / 8;
        static int gvar = -1;  // TA_GAE_SMALL (tuning knob for small rollouts): 0 off, 1 = 128-env CTAs x 16 chunks, 2 = 64-env CTAs, 3 = 32-env CTAs
        if (gvar < 0) { const char *e = getenv("TA_GAE_SMALL"); gvar = e ? atoi(e) : 3; if (gvar < 0 || gvar > 3) gvar = 3; }
        if (!gch && gvar && n / 128 < 2 * 148) {
            // small rollout (fewer than two 128-env CTAs per SM; BASELINE configs[3] is 128 x 16384): up to 16 chunks, so
            // that one pass covers 128 steps -- one DRAM round trip instead of two dependent ones
            const int chv = steps_chunks < 16 ? steps_chunks : 16;
            if (coop && stats3 && steps_chunks <= 16) {   // normalise in the same launch when every CTA is resident at once
                int dev = 0, sms = 0, occ = 0;
                CK(cudaGetDevice(&dev));
                CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gae_vec4_kernel<8, 16, 8, true>, 8 * chv, 0));
                const unsigned nb = blocks_for(n / 4, 8);
                if ((long long)nb <= (long long)occ * sms && nb <= TA_GAE_WORK_CTAS) {
                    gae_vec4_kernel<8, 16, 8, true><<<nb, dim3(8, chv), 0, st>>>(reward, v, v_next, last_v, done, gamma, lam, use_mask, T, n, adv_out,
                                                                            ret_out, stats3, coop);
                    if (coop_used) *coop_used = 1;
                    return launch_ok("gae_vec4_kernel (normalising)");
                }
            }
            if (gvar == 1)
                gae_vec4_kernel<8, 16, 32><<<blocks_for(n / 4, 32), dim3(32, chv), 0, st>>>(reward, v, v_next, last_v, done, gamma, lam, use_mask,
                                                                                            T, n, adv_out, ret_out, stats3, nullptr);
            else if (gvar == 2)
                gae_vec4_kernel<8, 16, 16><<<blocks_for(n / 4, 16), dim3(16, chv), 0, st>>>(reward, v, v_next, last_v, done, gamma, lam, use_mask,
                                                                                            T, n, adv_out, ret_out, stats3, nullptr);
            else
                gae_vec4_kernel<8, 16, 8><<<blocks_for(n / 4, 8), dim3(8, chv), 0, st>>>(reward, v, v_next, last_v, done, gamma, lam, use_mask, T,
                                                                                          n, adv_out, ret_out, stats3, nullptr);
            return launch_ok("gae_vec4_kernel (small rollout)");
        }
        int want = gch ? gch : (n / 128 >= 4 * 148 ? 4 : 8);
        int chv = steps_chunks;
        if (chv > want) chv = want;
        gae_vec4_kernel<8, 16, 32><<<blocks_for(n / 4, 32), dim3(32, chv), 0, st>>>(reward, v, v_next, last_v, done, gamma, lam, use_mask, T, n,
                                                                                    adv_out, ret_out, stats3, nullptr);
        return launch_ok("gae_vec4_kernel");
    }
    int ch = (T + GAE_L - 1) / GAE_L;
    if (ch > GAE_CH) ch = GAE_CH;
    dim3 block(32, ch);
    gae_kernel<<<blocks_for(n, 32), block, 0, st>>>(reward, v, v_next, last_v, done, gamma, lam, use_mask, T, n, adv_out, ret_out);
    if (int rc = launch_ok("gae_kernel")) return rc;
    if (stats3) {  // the scalar kernel has no fused moments: the separate pass
        unsigned nb = blocks_for((long long)T * n, 256 * 16);
        if (nb > 148 * 8) nb = 148 * 8;
        adv_stats_kernel<<<nb, 256, 0, st>>>(adv_out, (long long)T * n, stats3);
        return launch_ok("adv_stats_kernel");
    }
    return TA_OK;
}

int ta_gae(const float *reward, const float *v, const float *v_next, const float *last_v, const uint8_t *done, float gamma,
           float lam, int use_mask, int T, int64_t n, float *adv_out, float *ret_out, void *stream) {
    return gae_launch(reward, v, v_next, last_v, done, gamma, lam, use_mask, T, n, adv_out, ret_out, nullptr, stream);
}

int ta_gae_stats(const float *reward, const float *v, const float *v_next, const float *last_v, const uint8_t *done, float gamma,
                 float lam, int use_mask, int T, int64_t n, float *adv_out, float *ret_out, double *stats3, void *stream) {
    if (!stats3) return TA_E_INVALID;
    return gae_launch(reward, v, v_next, last_v, done, gamma, lam, use_mask, T, n, adv_out, ret_out, stats3, stream);
}

int ta_gae_normalized(const float *reward, const float *v, const float *v_next, const float *last_v, const uint8_t *done,
                      float gamma, float lam, int use_mask, int T, int64_t n, float *adv_out, float *ret_out, double *work,
                      void *stream) {
    double *work5 = work;
    if (!work5) return TA_E_INVALID;
    // work = (sum, sum of squares, count, ticket, spare) zeroed by the launch, then 3 unused and two slots per CTA.  A small rollout (one pass over T, every CTA of the
    // grid resident at once -- 128 x 16384 is) normalises inside the GAE launch; anything else takes the moments from the
    // GAE launch and one ta_adv_normalize pass.
    int used = 0;
    if (int rc = gae_launch(reward, v, v_next, last_v, done, gamma, lam, use_mask, T, n, adv_out, ret_out, work5, stream,
                            reinterpret_cast<unsigned int *>(work5 + 3), &used))
        return rc;
    if (used) return TA_OK;
    return ta_adv_normalize(adv_out, (int64_t)T * n, work5, stream);
}

static int pred_grid(const void *kern, int smem, int64_t M, unsigned *grid) {
    int dev = 0, sms = 0, occ = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, PR_THREADS, smem));
    long long g = (long long)sms * (occ > 0 ? occ : 1);
    *grid = (unsigned)(M < g ? M : g);
    return TA_OK;
}

int ta_pred_encoder(const void *x, int x_dtype, int64_t M, const float *w1, const float *s1, const float *t1, const float *w2,
                    const float *s2, const float *t2, const float *w3, const float *s3, const float *t3, void *z_bf16, void *stream) {
    if (!x || !z_bf16 || M <= 0 || !w1 || !s1 || !t1 || !w2 || !s2 || !t2 || !w3 || !s3 || !t3) return TA_E_INVALID;
    if (x_dtype != TA_STACK_U8 && x_dtype != TA_STACK_F32) return TA_E_INVALID;
    if ((((uintptr_t)w2 | (uintptr_t)w3) & 7u) != 0) return TA_E_INVALID;
    const PredEncArgs a{w1, s1, t1, w2, s2, t2, w3, s3, t3};
    unsigned g = 1;
    if (x_dtype == TA_STACK_U8) {
        if (int rc = pred_grid((const void *)pred_encoder_kernel<true>, PR_ENC_SMEM, M, &g)) return rc;
        pred_encoder_kernel<true><<<g, PR_THREADS, PR_ENC_SMEM, (cudaStream_t)stream>>>(x, a, (__nv_bfloat16 *)z_bf16, M);
    } else {
        if (int rc = pred_grid((const void *)pred_encoder_kernel<false>, PR_ENC_SMEM, M, &g)) return rc;
        pred_encoder_kernel<false><<<g, PR_THREADS, PR_ENC_SMEM, (cudaStream_t)stream>>>(x, a, (__nv_bfloat16 *)z_bf16, M);
    }
    return launch_ok("pred_encoder_kernel");
}

int ta_pred_decoder(const void *z_bf16, int64_t M, const float *w1, const float *b1, const float *w2, const float *b2, const float *w3,
                    float b3, float *out, void *stream) {
    if (!z_bf16 || !out || M <= 0 || !w1 || !b1 || !w2 || !b2 || !w3) return TA_E_INVALID;
    if ((((uintptr_t)w1 | (uintptr_t)w2) & 15u) != 0) return TA_E_INVALID;
    const PredDecArgs a{w1, b1, w2, b2, w3, b3};
    unsigned g = 1;
    if (int rc = pred_grid((const void *)pred_decoder_kernel, PR_DEC_SMEM, M, &g)) return rc;
    pred_decoder_kernel<<<g, PR_THREADS, PR_DEC_SMEM, (cudaStream_t)stream>>>((const __nv_bfloat16 *)z_bf16, a, out, M);
    return launch_ok("pred_decoder_kernel");
}

int ta_lstm_gates(const float *gx, const float *gh, const float *bias, float *c, void *h_out_bf16, int64_t ld_h, int64_t B, int H,
                  void *stream) {
    if (!gx || !bias || !c || !h_out_bf16 || B <= 0 || H <= 0 || (H & 3) || ld_h < H || (ld_h & 3)) return TA_E_INVALID;
    if ((((uintptr_t)gx | (uintptr_t)gh | (uintptr_t)bias | (uintptr_t)c) & 15u) || ((uintptr_t)h_out_bf16 & 7u)) return TA_E_INVALID;
    lstm_gates_kernel<<<blocks_for(B * (H / 4), 256), 256, 0, (cudaStream_t)stream>>>(gx, gh, bias, c, (__nv_bfloat16 *)h_out_bf16, ld_h, B, H);
    return launch_ok("lstm_gates_kernel");
}

int ta_adv_stats(const float *adv, int64_t count, double *stats3, void *stream) {
    if (!adv || !stats3 || count <= 0) return TA_E_INVALID;
    CK(cudaMemsetAsync(stats3, 0, 3 * sizeof(double), (cudaStream_t)stream));
    unsigned nb = blocks_for(count, 256 * 16);
    if (nb > 148 * 8) nb = 148 * 8;
    adv_stats_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(adv, count, stats3);
    return launch_ok("adv_stats_kernel");
}

int ta_adv_normalize(float *adv, int64_t count, const double *stats3, void *stream) {
    if (!adv || !stats3 || count <= 0) return TA_E_INVALID;
    unsigned nb = blocks_for(count, 256 * 16);
    if (nb > 148 * 8) nb = 148 * 8;
    adv_normalize_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(adv, count, stats3);
    return launch_ok("adv_normalize_kernel");
}

int ta_her_plan(const float *p, const uint8_t *done, int T, int64_t n, int first_record, uint64_t seed, uint64_t env_id0,
                const uint8_t *chosen_in, uint8_t *uniq_out, uint8_t *m_out, uint16_t *plan_out, void *stream) {
    if (!p || !done || !plan_out || T <= 0 || n <= 0 || first_record < 0 || first_record >= HER_MAXLEN) return TA_E_INVALID;
    CK(cudaMemsetAsync(plan_out, 0xFF, (size_t)T * n * 4 * sizeof(uint16_t), (cudaStream_t)stream));
    if (uniq_out) CK(cudaMemsetAsync(uniq_out, 0xFF, (size_t)T * n * HER_MAXLEN, (cudaStream_t)stream));
    if (m_out) CK(cudaMemsetAsync(m_out, 0, (size_t)T * n, (cudaStream_t)stream));
    her_plan_kernel<<<blocks_for(n, HER_WARPS), 32 * HER_WARPS, 0, (cudaStream_t)stream>>>(
        p, done, T, n, first_record, (uint32_t)seed, (uint32_t)(seed >> 32), env_id0, chosen_in, uniq_out, m_out, plan_out);
    return launch_ok("her_plan_kernel");
}

}  // extern "C"

namespace {
// Everything the handle-free entry points cache is kept PER DEVICE (one process may drive several GPUs): the failure
// flag of the tcgen05 kernels (set when a bounded MMA-barrier wait gave up), SM counts, CTAs per SM and the kernels'
// shared-memory attributes, which are per-device state of the CUDA runtime.
constexpr int MAX_DEV = 64;
int *g_tc_fail_dev[MAX_DEV] = {};
int cur_dev() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    return d & (MAX_DEV - 1);
}
int tc_fail_flag(int **out) {   // the current device's flag, allocated on first use
    int *&f = g_tc_fail_dev[cur_dev()];
    if (!f) {
        CK(cudaMalloc(&f, sizeof(int)));
        CK(cudaMemset(f, 0, sizeof(int)));
    }
    *out = f;
    return TA_OK;
}
int g_use_tc = -1;         // -1: read TA_CONV1_TC on first use (default 1)

int g_bwd_tc = -1;         // conv1 weight gradient on tcgen05: -1 = read TA_CONV1_BWD_TC on first use

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda); a rank-3 bf16 tensor with a 128-byte-swizzled box
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int tensor_map_3d_bf16(CUtensorMap *map, void *base, const cuuint64_t dims[3], const cuuint64_t strides[2], const cuuint32_t box[3]) {
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) return cuda_fail(cudaErrorNotSupported, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (encode_fn)fn;
    }
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
    return TA_OK;
}

// Resident CTAs per SM of a TMEM-allocating kernel from its own footprint (the occupancy API answers 1 for such
// kernels): registers, shared memory with the large carve-out, and 512 TMEM columns per SM.
int tc_ctas_per_sm(const void *kern, int threads, int dyn_smem, int tmem_cols, int *out) {
    int dev = 0, regs_sm = 0, smem_sm = 0, smem_rsv = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
    CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    CK(cudaDeviceGetAttribute(&smem_rsv, cudaDevAttrReservedSharedMemoryPerBlock, dev));
    cudaFuncAttributes fa;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (dyn_smem > 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
    CK(cudaFuncGetAttributes(&fa, kern));
    const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * threads, smem_cta = (int)fa.sharedSizeBytes + dyn_smem + smem_rsv;
    int n = 512 / tmem_cols;
    if (regs_sm / regs_cta < n) n = regs_sm / regs_cta;
    if (smem_sm / smem_cta < n) n = smem_sm / smem_cta;
    *out = n > 0 ? n : 1;
    return TA_OK;
}

template <typename K>
int conv1_grid(K kern, long long batch, int *grid) {
    int dev = 0, sms = 0, occ = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C1_THREADS, 0));
    long long g = (long long)sms * (occ > 0 ? occ : 1);
    *grid = (int)(batch < g ? batch : g);
    return TA_OK;
}
}  // namespace

extern "C" {

}  // extern "C"

namespace {
int conv1_fwd_impl(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch, void *y_bf16,
                   uint32_t *relu_mask, void *stream);
}

extern "C" {

int ta_conv1_fwd(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch, void *y_bf16,
                 void *stream) {
    return conv1_fwd_impl(x, x_dtype, x_stride, w4, b4, batch, y_bf16, nullptr, stream);
}

int ta_conv1_fwd_mask(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch, void *y_bf16,
                      uint32_t *relu_mask, void *stream) {
    if (!relu_mask || ((uintptr_t)relu_mask & 3u) || batch * NCELL >= (1ll << 31)) return TA_E_INVALID;
    return conv1_fwd_impl(x, x_dtype, x_stride, w4, b4, batch, y_bf16, relu_mask, stream);
}

}  // extern "C"

namespace {
int conv1_fwd_impl(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch, void *y_bf16,
                   uint32_t *relu_mask, void *stream) {
    if (!x || !w4 || !b4 || !y_bf16 || batch <= 0 || x_stride < 4 * NCELL || (x_dtype != TA_X_F32 && x_dtype != TA_X_U8))
        return TA_E_INVALID;
    if ((uintptr_t)y_bf16 & 7u) return TA_E_INVALID;
    if (g_use_tc < 0) { const char *e = getenv("TA_CONV1_TC"); g_use_tc = e ? atoi(e) : 2; }
    if (g_use_tc >= 2 && !((uintptr_t)y_bf16 & 15u) && batch * NCELL < (1ll << 31)) {
        // default: the warp-specialised kernel with tensor-map stores (ta_conv1_fwd_ws.cuh; 135 / 156 us without / with the bit
        // mask per 4096 samples); TA_CONV1_TC=1 selects conv1_fwd_tc_kernel (148 / 164 us, bit-identical), 0 the FP32-FMA kernel
        static int sms_d[MAX_DEV] = {};
        int &sms = sms_d[cur_dev()];
        if (!sms) {
            int dev = 0;
            CK(cudaGetDevice(&dev));
            CK(cudaFuncSetAttribute(conv1_fwd_ws_kernel<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM));
            CK(cudaFuncSetAttribute(conv1_fwd_ws_kernel<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM));
            CK(cudaFuncSetAttribute(conv1_fwd_ws_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM));
            CK(cudaFuncSetAttribute(conv1_fwd_ws_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM));
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        int *g_tc_fail = nullptr;
        if (int rc = tc_fail_flag(&g_tc_fail)) return rc;
        // y as [batch * 33 image rows][33 pixels][64 channels]; one store = one image row
        const cuuint64_t dims[3] = {64, (cuuint64_t)C1_OUT, (cuuint64_t)batch * C1_OUT}, strides[2] = {128, (cuuint64_t)C1_OUT * 128};
        const cuuint32_t box[3] = {64, (cuuint32_t)C1_OUT, 1};
        CUtensorMap map;
        if (int rc = tensor_map_3d_bf16(&map, y_bf16, dims, strides, box)) return rc;
        const long long ntiles = (batch * GS + FW_ROWS - 1) / FW_ROWS;
        const int g = (int)(ntiles < sms ? ntiles : sms);   // one persistent CTA per SM (all 512 TMEM columns)
        cudaStream_t st = (cudaStream_t)stream;
        static int dbg = -1;
        if (dbg < 0) { const char *e = getenv("TA_FW_DBG"); dbg = e ? atoi(e) : 0; }
        if (x_dtype == TA_X_U8 && relu_mask)
            conv1_fwd_ws_kernel<uint8_t, true><<<g, FW_THREADS, FW_SMEM, st>>>((const uint8_t *)x, x_stride, w4, b4, batch, map, relu_mask, g_tc_fail, dbg);
        else if (x_dtype == TA_X_U8)
            conv1_fwd_ws_kernel<uint8_t, false><<<g, FW_THREADS, FW_SMEM, st>>>((const uint8_t *)x, x_stride, w4, b4, batch, map, nullptr, g_tc_fail, dbg);
        else if (relu_mask)
            conv1_fwd_ws_kernel<float, true><<<g, FW_THREADS, FW_SMEM, st>>>((const float *)x, x_stride, w4, b4, batch, map, relu_mask, g_tc_fail, dbg);
        else
            conv1_fwd_ws_kernel<float, false><<<g, FW_THREADS, FW_SMEM, st>>>((const float *)x, x_stride, w4, b4, batch, map, nullptr, g_tc_fail, dbg);
        return launch_ok("conv1_fwd_ws_kernel");
    }
    if (g_use_tc || relu_mask) {  // tcgen05 version of the layer (the only one that writes the ReLU bit mask) (TA_CONV1_TC=0 / ta_debug_conv1_tc(0) select the FP32-FMA kernel)
        static int per_sm_u8_d[MAX_DEV] = {}, per_sm_f32_d[MAX_DEV] = {}, sms_d[MAX_DEV] = {};
        const int cd = cur_dev();
        int &per_sm_u8 = per_sm_u8_d[cd], &per_sm_f32 = per_sm_f32_d[cd], &sms = sms_d[cd];
        if (!sms) {
            int dev = 0;
            CK(cudaGetDevice(&dev));
            int a = 0, b = 0;
            if (int rc = tc_ctas_per_sm((const void *)conv1_fwd_tc_kernel<uint8_t, false>, TC_THREADS, 0, TC_NT, &a)) return rc;
            if (int rc = tc_ctas_per_sm((const void *)conv1_fwd_tc_kernel<uint8_t, true>, TC_THREADS, 0, TC_NT, &b)) return rc;
            per_sm_u8 = a < b ? a : b;
            if (int rc = tc_ctas_per_sm((const void *)conv1_fwd_tc_kernel<float, false>, TC_THREADS, 0, TC_NT, &a)) return rc;
            if (int rc = tc_ctas_per_sm((const void *)conv1_fwd_tc_kernel<float, true>, TC_THREADS, 0, TC_NT, &b)) return rc;
            per_sm_f32 = a < b ? a : b;
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            if (getenv("TA_VERBOSE")) fprintf(stderr, "conv1_tc: %d / %d CTAs per SM (u8 / f32 input), %d SMs\n", per_sm_u8, per_sm_f32, sms);
        }
        int *g_tc_fail = nullptr;
        if (int rc = tc_fail_flag(&g_tc_fail)) return rc;
        const int per_sm = x_dtype == TA_X_U8 ? per_sm_u8 : per_sm_f32;
        const long long ntiles = (batch * NCELL + TC_M - 1) / TC_M;
        const long long cap = (long long)sms * per_sm;
        const int g = (int)(ntiles < cap ? ntiles : cap);
        cudaStream_t st = (cudaStream_t)stream;
        __nv_bfloat16 *yb = (__nv_bfloat16 *)y_bf16;
        if (x_dtype == TA_X_U8 && relu_mask)
            conv1_fwd_tc_kernel<uint8_t, true><<<g, TC_THREADS, 0, st>>>((const uint8_t *)x, x_stride, w4, b4, batch, yb, relu_mask, g_tc_fail);
        else if (x_dtype == TA_X_U8)
            conv1_fwd_tc_kernel<uint8_t, false><<<g, TC_THREADS, 0, st>>>((const uint8_t *)x, x_stride, w4, b4, batch, yb, nullptr, g_tc_fail);
        else if (relu_mask)
            conv1_fwd_tc_kernel<float, true><<<g, TC_THREADS, 0, st>>>((const float *)x, x_stride, w4, b4, batch, yb, relu_mask, g_tc_fail);
        else
            conv1_fwd_tc_kernel<float, false><<<g, TC_THREADS, 0, st>>>((const float *)x, x_stride, w4, b4, batch, yb, nullptr, g_tc_fail);
        return launch_ok("conv1_fwd_tc_kernel");
    }
    int grid = 1;
    if (x_dtype == TA_X_U8) {
        if (int rc = conv1_grid(conv1_fwd_kernel<uint8_t>, batch, &grid)) return rc;
        conv1_fwd_kernel<uint8_t><<<grid, C1_THREADS, 0, (cudaStream_t)stream>>>((const uint8_t *)x, x_stride, w4, b4, batch,
                                                                                (__nv_bfloat16 *)y_bf16);
    } else {
        if (int rc = conv1_grid(conv1_fwd_kernel<float>, batch, &grid)) return rc;
        conv1_fwd_kernel<float><<<grid, C1_THREADS, 0, (cudaStream_t)stream>>>((const float *)x, x_stride, w4, b4, batch,
                                                                              (__nv_bfloat16 *)y_bf16);
    }
    return launch_ok("conv1_fwd_kernel");
}

// the tcgen05 weight-gradient kernel; planes == nullptr: dy is one channels-last tensor; relu_mask (with planes): y is not read
int launch_conv1_bwd_tc(const void *x, int x_dtype, int64_t x_stride, const void *y_bf16, const void *dy_bf16, const void *planes,
                        const void *relu_mask, int64_t batch, float *dw4, float *db4, void *stream, int class_major = 0) {
    static int per_sm_d[MAX_DEV][6] = {}, sms_d[MAX_DEV] = {};
    const int cd = cur_dev();
    int *per_sm = per_sm_d[cd], &sms = sms_d[cd];
    const int dyn = TCB_A_BYTES + 2 * TCB_B_BYTES;
    if (!sms) {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        if (int rc = tc_ctas_per_sm((const void *)conv1_bwd_tc_kernel<uint8_t, false, false>, TC_THREADS, dyn, TCB_COLS, &per_sm[0])) return rc;
        if (int rc = tc_ctas_per_sm((const void *)conv1_bwd_tc_kernel<float, false, false>, TC_THREADS, dyn, TCB_COLS, &per_sm[1])) return rc;
        if (int rc = tc_ctas_per_sm((const void *)conv1_bwd_tc_kernel<uint8_t, true, false>, TC_THREADS, dyn, TCB_COLS, &per_sm[2])) return rc;
        if (int rc = tc_ctas_per_sm((const void *)conv1_bwd_tc_kernel<float, true, false>, TC_THREADS, dyn, TCB_COLS, &per_sm[3])) return rc;
        if (int rc = tc_ctas_per_sm((const void *)conv1_bwd_tc_kernel<uint8_t, true, true>, TC_THREADS, dyn, TCB_COLS, &per_sm[4])) return rc;
        if (int rc = tc_ctas_per_sm((const void *)conv1_bwd_tc_kernel<float, true, true>, TC_THREADS, dyn, TCB_COLS, &per_sm[5])) return rc;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (getenv("TA_VERBOSE"))
            fprintf(stderr, "conv1_bwd_tc: %d / %d / %d CTAs per SM (dense / planes / planes + bit mask), %d SMs\n", per_sm[0], per_sm[2], per_sm[4], sms);
    }
    int *g_tc_fail = nullptr;
    if (int rc = tc_fail_flag(&g_tc_fail)) return rc;
    if (batch * NCELL >= (1ll << 31)) return TA_E_INVALID;
    const long long ntiles = (batch * NCELL + TC_M - 1) / TC_M;
    const int variant = (planes ? (relu_mask ? 4 : 2) : 0) + (x_dtype == TA_X_U8 ? 0 : 1);
    const long long cap = (long long)sms * per_sm[variant];
    const int g = (int)(ntiles < cap ? ntiles : cap);
    const __nv_bfloat16 *pl = (const __nv_bfloat16 *)planes;
    const __nv_bfloat16 *yb = (const __nv_bfloat16 *)y_bf16, *dyb = (const __nv_bfloat16 *)dy_bf16;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t *mk = (const uint32_t *)relu_mask;
    const long long pl_pos = class_major ? 64 : 256, pl_cls = class_major ? batch * NCELL * 64 : 64;
#define TA_BWD_ARGS(XT) (const XT *)x, x_stride, yb, dyb, pl, mk, batch, dw4, db4, 0u, g_tc_fail, pl_pos, pl_cls
    switch (variant) {
        case 0: conv1_bwd_tc_kernel<uint8_t, false, false><<<g, TC_THREADS, dyn, st>>>(TA_BWD_ARGS(uint8_t)); break;
        case 1: conv1_bwd_tc_kernel<float, false, false><<<g, TC_THREADS, dyn, st>>>(TA_BWD_ARGS(float)); break;
        case 2: conv1_bwd_tc_kernel<uint8_t, true, false><<<g, TC_THREADS, dyn, st>>>(TA_BWD_ARGS(uint8_t)); break;
        case 3: conv1_bwd_tc_kernel<float, true, false><<<g, TC_THREADS, dyn, st>>>(TA_BWD_ARGS(float)); break;
        case 4: conv1_bwd_tc_kernel<uint8_t, true, true><<<g, TC_THREADS, dyn, st>>>(TA_BWD_ARGS(uint8_t)); break;
        default: conv1_bwd_tc_kernel<float, true, true><<<g, TC_THREADS, dyn, st>>>(TA_BWD_ARGS(float)); break;
    }
#undef TA_BWD_ARGS
    return launch_ok("conv1_bwd_tc_kernel");
}
}  // namespace

extern "C" {

int ta_conv1_fwd_add(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch,
                     const void *addend_bf16, int relu, void *y_bf16, void *stream) {
    if (!x || !w4 || !b4 || !y_bf16 || batch <= 0 || x_stride < 4 * NCELL || (x_dtype != TA_X_F32 && x_dtype != TA_X_U8)) return TA_E_INVALID;
    if (((uintptr_t)y_bf16 | (uintptr_t)addend_bf16) & 7u) return TA_E_INVALID;
    int grid = 1;
    if (x_dtype == TA_X_U8) {
        if (int rc = conv1_grid(conv1_fwd_kernel<uint8_t>, batch, &grid)) return rc;
        conv1_fwd_kernel<uint8_t><<<grid, C1_THREADS, 0, (cudaStream_t)stream>>>((const uint8_t *)x, x_stride, w4, b4, batch, (__nv_bfloat16 *)y_bf16,
                                                                                 (const __nv_bfloat16 *)addend_bf16, relu);
    } else {
        if (int rc = conv1_grid(conv1_fwd_kernel<float>, batch, &grid)) return rc;
        conv1_fwd_kernel<float><<<grid, C1_THREADS, 0, (cudaStream_t)stream>>>((const float *)x, x_stride, w4, b4, batch, (__nv_bfloat16 *)y_bf16,
                                                                               (const __nv_bfloat16 *)addend_bf16, relu);
    }
    return launch_ok("conv1_fwd_kernel (add)");
}

int ta_conv1_bwd_planes(const void *x, int x_dtype, int64_t x_stride, const void *y_bf16, const uint32_t *relu_mask,
                        const void *planes_bf16, int class_major, int64_t batch, float *dw4, float *db4, void *stream) {
    if (!x || (!y_bf16 && !relu_mask) || !planes_bf16 || !dw4 || !db4 || batch <= 0 || x_stride < 4 * NCELL ||
        (x_dtype != TA_X_F32 && x_dtype != TA_X_U8))
        return TA_E_INVALID;
    if (((uintptr_t)y_bf16 | (uintptr_t)planes_bf16) & 15u) return TA_E_INVALID;
    CK(cudaMemsetAsync(dw4, 0, 256 * 16 * sizeof(float), (cudaStream_t)stream));
    CK(cudaMemsetAsync(db4, 0, 256 * sizeof(float), (cudaStream_t)stream));
    return launch_conv1_bwd_tc(x, x_dtype, x_stride, y_bf16, nullptr, planes_bf16, relu_mask, batch, dw4, db4, stream, class_major != 0);
}

int ta_conv1_bwd(const void *x, int x_dtype, int64_t x_stride, const void *y_bf16, const void *dy_bf16, int64_t batch,
                 float *dw4, float *db4, void *stream) {
    if (!x || !y_bf16 || !dy_bf16 || !dw4 || !db4 || batch <= 0 || x_stride < 4 * NCELL ||
        (x_dtype != TA_X_F32 && x_dtype != TA_X_U8))
        return TA_E_INVALID;
    if (((uintptr_t)y_bf16 | (uintptr_t)dy_bf16) & 7u) return TA_E_INVALID;
    CK(cudaMemsetAsync(dw4, 0, 256 * 16 * sizeof(float), (cudaStream_t)stream));
    CK(cudaMemsetAsync(db4, 0, 256 * sizeof(float), (cudaStream_t)stream));
    if (g_bwd_tc < 0) { const char *e = getenv("TA_CONV1_BWD_TC"); g_bwd_tc = e ? atoi(e) != 0 : 1; }
    if (g_bwd_tc && !(((uintptr_t)y_bf16 | (uintptr_t)dy_bf16) & 15u) && batch * NCELL < (1ll << 31))
        return launch_conv1_bwd_tc(x, x_dtype, x_stride, y_bf16, dy_bf16, nullptr, nullptr, batch, dw4, db4, stream);
    int grid = 1;
    if (x_dtype == TA_X_U8) {
        if (int rc = conv1_grid(conv1_bwd_kernel<uint8_t>, batch, &grid)) return rc;
        conv1_bwd_kernel<uint8_t><<<grid, C1_THREADS, 0, (cudaStream_t)stream>>>(
            (const uint8_t *)x, x_stride, (const __nv_bfloat16 *)y_bf16, (const __nv_bfloat16 *)dy_bf16, batch, dw4, db4);
    } else {
        if (int rc = conv1_grid(conv1_bwd_kernel<float>, batch, &grid)) return rc;
        conv1_bwd_kernel<float><<<grid, C1_THREADS, 0, (cudaStream_t)stream>>>(
            (const float *)x, x_stride, (const __nv_bfloat16 *)y_bf16, (const __nv_bfloat16 *)dy_bf16, batch, dw4, db4);
    }
    return launch_ok("conv1_bwd_kernel");
}

/* development probe (not part of the ABI): device buffer of 16 int64 that CTA 0 of conv2_dgrad_planes_ws_kernel (8 values) /
 * conv2_dgrad_conv1_wgrad_kernel (10 values) fills with the cycles its roles spend waiting; NULL switches it off */
static long long *g_dgrad_prof = nullptr;
int ta_debug_dgrad_profile(long long *prof8) {
    g_dgrad_prof = prof8;
    return TA_OK;
}

int ta_conv2_dgrad_prep(const void *w_bf16, int64_t stride_o, int64_t stride_i, int64_t stride_y, int64_t stride_x, void *wimg_bf16,
                        void *stream) {
    if (!w_bf16 || !wimg_bf16 || ((uintptr_t)wimg_bf16 & 15u)) return TA_E_INVALID;
    conv2_dgrad_prep_kernel<<<36, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)w_bf16, stride_o, stride_i, stride_y, stride_x,
                                                                 (__nv_bfloat16 *)wimg_bf16);
    return launch_ok("conv2_dgrad_prep_kernel");
}

int ta_conv2_dgrad_planes(const void *dz_bf16, const void *wimg_bf16, const uint32_t *relu_mask, int64_t batch, int class_major,
                          void *planes_bf16, void *stream) {
    if (!dz_bf16 || !wimg_bf16 || !planes_bf16 || batch <= 0 || batch * (DG_P * DG_P) >= (1ll << 31)) return TA_E_INVALID;
    if (((uintptr_t)dz_bf16 | (uintptr_t)wimg_bf16 | (uintptr_t)planes_bf16) & 15u) return TA_E_INVALID;
    static int sms_d[MAX_DEV] = {};
    int &sms = sms_d[cur_dev()];
    const int ws = class_major != 0;   // class-major output: the warp-specialised kernel; position-major: the single-role one
    if (!sms) {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaFuncSetAttribute(conv2_dgrad_planes_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM));
        CK(cudaFuncSetAttribute(conv2_dgrad_planes_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DGW_SMEM));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int *g_tc_fail = nullptr;
    if (int rc = tc_fail_flag(&g_tc_fail)) return rc;
    const long long ntiles = (batch * (DG_P * DG_P) + TC_M - 1) / TC_M;
    const int g = (int)(ntiles < sms ? ntiles : sms);   // one persistent CTA per SM (208 / 216 KB of shared memory each)
    if (ws) {
        // the class-major planes [4][npos][64] as a rank-3 tensor; the kernel stores boxes of 32 positions x 64 channels of one
        // class from 128-byte-swizzled shared memory
        const cuuint64_t npos = (cuuint64_t)batch * (DG_P * DG_P);
        const cuuint64_t dims[3] = {64, npos, 4}, strides[2] = {128, npos * 128};
        const cuuint32_t box[3] = {64, 32, 1};
        CUtensorMap map;
        if (int rc = tensor_map_3d_bf16(&map, planes_bf16, dims, strides, box)) return rc;
        conv2_dgrad_planes_ws_kernel<<<g, DGW_THREADS, DGW_SMEM, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dz_bf16, (const uint4 *)wimg_bf16,
                                                                                       relu_mask, batch, map, g_tc_fail, g_dgrad_prof);
        return launch_ok("conv2_dgrad_planes_ws_kernel");
    }
    conv2_dgrad_planes_tc_kernel<<<g, DG_THREADS, DG_SMEM, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dz_bf16, (const uint4 *)wimg_bf16, relu_mask,
                                                                                  batch, (__nv_bfloat16 *)planes_bf16, g_tc_fail);
    return launch_ok("conv2_dgrad_planes_tc_kernel");
}

int ta_conv2_dgrad_conv1_bwd(const void *dz_bf16, const void *wimg_bf16, const uint32_t *relu_mask, const void *x, int x_dtype,
                             int64_t x_stride, int64_t batch, float *dw4, float *db4, void *stream) {
    if (!dz_bf16 || !wimg_bf16 || !relu_mask || !x || !dw4 || !db4 || batch <= 0 || batch * (DG_P * DG_P) >= (1ll << 31) ||
        x_stride < 4 * NCELL || (x_dtype != TA_X_F32 && x_dtype != TA_X_U8))
        return TA_E_INVALID;
    if ((((uintptr_t)dz_bf16 | (uintptr_t)wimg_bf16 | (uintptr_t)relu_mask) & 15u)) return TA_E_INVALID;
    static int sms_d[MAX_DEV] = {};
    int &sms = sms_d[cur_dev()];
    if (!sms) {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaFuncSetAttribute(conv2_dgrad_conv1_wgrad_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM));
        CK(cudaFuncSetAttribute(conv2_dgrad_conv1_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int *g_tc_fail = nullptr;
    if (int rc = tc_fail_flag(&g_tc_fail)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(dw4, 0, 256 * 16 * sizeof(float), st));
    CK(cudaMemsetAsync(db4, 0, 256 * sizeof(float), st));
    const long long ntiles = (batch * (DG_P * DG_P) + TC_M - 1) / TC_M;
    const int g = (int)(ntiles < sms ? ntiles : sms);   // one persistent CTA per SM (221 KB of shared memory, all 512 TMEM columns)
    if (x_dtype == TA_X_U8)
        conv2_dgrad_conv1_wgrad_kernel<uint8_t><<<g, SB_THREADS, SB_SMEM, st>>>((const __nv_bfloat16 *)dz_bf16, (const uint4 *)wimg_bf16, relu_mask,
                                                                               (const uint8_t *)x, x_stride, batch, dw4, db4, g_tc_fail, g_dgrad_prof);
    else
        conv2_dgrad_conv1_wgrad_kernel<float><<<g, SB_THREADS, SB_SMEM, st>>>((const __nv_bfloat16 *)dz_bf16, (const uint4 *)wimg_bf16, relu_mask,
                                                                             (const float *)x, x_stride, batch, dw4, db4, g_tc_fail, g_dgrad_prof);
    return launch_ok("conv2_dgrad_conv1_wgrad_kernel");
}

int ta_im2col_s2(const void *x_bf16, void *cols_bf16, int64_t batch, int H, int W, int C, int ksize, void *stream) {
    if (!x_bf16 || !cols_bf16 || batch <= 0 || H < ksize || W < ksize || C <= 0 || (C & 7) || ksize < 1 || ksize > 8) return TA_E_INVALID;
    if (((uintptr_t)x_bf16 | (uintptr_t)cols_bf16) & 15u) return TA_E_INVALID;
    const int OH = (H - ksize) / 2 + 1, OW = (W - ksize) / 2 + 1;
    const long long total = (long long)batch * OH * OW * ksize * ksize * (C >> 3);
    if (total >= (1ll << 31)) return TA_E_INVALID;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    im2col_s2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)x_bf16, (uint4 *)cols_bf16, (unsigned)total, H, W, C >> 3,
                                                                      OH, OW, ksize);
    return launch_ok("im2col_s2_kernel");
}

int ta_parity_class_weights(const void *w_bf16, int64_t stride_o, int64_t stride_i, int64_t stride_y, int64_t stride_x, int cout, int cin,
                            int ksize, void *out_bf16, void *stream) {
    if (!w_bf16 || !out_bf16 || cout <= 0 || cin <= 0 || (ksize != 3 && ksize != 4)) return TA_E_INVALID;
    const long long total = 16ll * cout * cin;
    if (total >= (1ll << 31)) return TA_E_INVALID;
    parity_class_weights_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)w_bf16, stride_o, stride_i,
                                                                                      stride_y, stride_x, cout, cin, ksize,
                                                                                      (__nv_bfloat16 *)out_bf16);
    return launch_ok("parity_class_weights_kernel");
}

int ta_planes_to_dense_relu(const void *planes_bf16, const void *y_bf16, void *dz_bf16, int64_t batch, int H, int W, int C, int ksize,
                            void *stream) {
    if (!planes_bf16 || !y_bf16 || !dz_bf16 || batch <= 0 || (ksize != 3 && ksize != 4) || H < ksize || W < ksize || C <= 0 || (C & 7))
        return TA_E_INVALID;
    if (((uintptr_t)planes_bf16 | (uintptr_t)y_bf16 | (uintptr_t)dz_bf16) & 15u) return TA_E_INVALID;
    if ((H - ksize) / 2 + 1 < (H - 1) / 2 || (W - ksize) / 2 + 1 < (W - 1) / 2) return TA_E_INVALID;  // every pixel has a plane entry
    const long long total = (long long)batch * H * W * (C >> 3);
    if (total >= (1ll << 31)) return TA_E_INVALID;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    planes_to_dense_relu_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)planes_bf16, (const uint4 *)y_bf16,
                                                                                 (uint4 *)dz_bf16, (unsigned)total, H, W, C >> 3,
                                                                                 (H - ksize) / 2 + 1, (W - ksize) / 2 + 1);
    return launch_ok("planes_to_dense_relu_kernel");
}

int ta_col2im_s2(const void *dcols_bf16, void *dx_bf16, int64_t batch, int H, int W, int C, int ksize, void *stream) {
    if (!dcols_bf16 || !dx_bf16 || batch <= 0 || H < ksize || W < ksize || C <= 0 || (C & 7) || (ksize != 3 && ksize != 4))
        return TA_E_INVALID;
    if (((uintptr_t)dcols_bf16 | (uintptr_t)dx_bf16) & 15u) return TA_E_INVALID;
    const int OH = (H - ksize) / 2 + 1, OW = (W - ksize) / 2 + 1;
    if (batch * H > 0x7FFFFFFFll) return TA_E_INVALID;
    const unsigned nb = (unsigned)(batch * H);
    if (ksize == 3)
        col2im_s2_kernel<3><<<nb, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dcols_bf16, (__nv_bfloat16 *)dx_bf16, H, W, C,
                                                                OH, OW);
    else
        col2im_s2_kernel<4><<<nb, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dcols_bf16, (__nv_bfloat16 *)dx_bf16, H, W, C,
                                                                OH, OW);
    return launch_ok("col2im_s2_kernel");
}

int ta_channel_sum_bf16(const void *x_bf16, int64_t rows, int C, float *out, void *stream) {
    if (!x_bf16 || !out || rows <= 0 || C < 64 || C > 2048 || (C & (C - 1)) || ((uintptr_t)x_bf16 & 15u)) return TA_E_INVALID;
    CK(cudaMemsetAsync(out, 0, C * sizeof(float), (cudaStream_t)stream));
    const int nrl = 256 / (C / 8);
    unsigned nb = blocks_for(rows, (long long)nrl * 16);
    if (nb > 148u * 8u) nb = 148u * 8u;
    channel_sum_bf16_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x_bf16, rows, C, out);
    return launch_ok("channel_sum_bf16_kernel");
}

/* ---- the small kernels of the hand-scheduled optimiser step (ta_train.cuh) ---- */
static_assert(sizeof(ta_tinet_prep_args) == sizeof(PrepArgs), "public and device struct must match");
static_assert(sizeof(ta_tinet_grad_args) == sizeof(GradArgs), "public and device struct must match");

static long long relu_bwd_grid(long long rows) {   // CTAs: 32 rows each at least, at most 4 per SM
    long long g = (rows + 31) / 32;
    if (g > 4 * 148) g = 4 * 148;
    return g < 1 ? 1 : g;
}
int64_t ta_relu_bwd_bias_scratch_floats(int64_t rows, int C) { return relu_bwd_grid(rows) * C + 4; }

int ta_relu_bwd_bias(const void *dy_bf16, int64_t ld_dy, const void *y_bf16, void *dz_bf16, int64_t rows, int C, float *db_out,
                     float *scratch, void *stream) {
    if (!dy_bf16 || !db_out || !scratch || rows <= 0 || C < 64 || C > 2048 || (C & 7) || (256 % (C >> 3)) || ld_dy < C || (ld_dy & 7))
        return TA_E_INVALID;
    if ((y_bf16 == nullptr) != (dz_bf16 == nullptr)) return TA_E_INVALID;
    if (((uintptr_t)dy_bf16 | (uintptr_t)y_bf16 | (uintptr_t)dz_bf16 | (uintptr_t)db_out | (uintptr_t)scratch) & 15u) return TA_E_INVALID;
    const long long g = relu_bwd_grid(rows);
    relu_bwd_bias_kernel<<<(unsigned)g, RB_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dy_bf16, ld_dy, (const __nv_bfloat16 *)y_bf16,
                                                                             (__nv_bfloat16 *)dz_bf16, rows, C, db_out, scratch, 0, 0);
    return launch_ok("relu_bwd_bias_kernel");
}

int ta_planes_relu_bwd_bias(const void *planes_bf16, const void *y_bf16, void *dz_bf16, int64_t batch, int H, int W, int C, float *db_out,
                            float *scratch, void *stream) {
    if (!planes_bf16 || !y_bf16 || !dz_bf16 || !db_out || !scratch || batch <= 0 || H < 2 || W < 2 || C < 64 || C > 2048 ||
        (C & 7) || (256 % (C >> 3)))
        return TA_E_INVALID;
    if (((uintptr_t)planes_bf16 | (uintptr_t)y_bf16 | (uintptr_t)dz_bf16 | (uintptr_t)db_out | (uintptr_t)scratch) & 15u) return TA_E_INVALID;
    const long long rows = batch * H * W, g = relu_bwd_grid(rows);
    if (rows * 4 >= (1ll << 32)) return TA_E_INVALID;   // the kernel's plane addressing divides in 32 bits
    relu_bwd_bias_kernel<<<(unsigned)g, RB_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)planes_bf16, 0, (const __nv_bfloat16 *)y_bf16,
                                                                             (__nv_bfloat16 *)dz_bf16, rows, C, db_out, scratch, H, W);
    return launch_ok("relu_bwd_bias_kernel (planes)");
}

int ta_ppo_actor_loss(const void *logits_bf16, const int32_t *act, const float *old_logp, const float *adv, int B, float clip,
                      float ent_coef, void *dlogits_bf16, float *loss_out, float *db_head, float *step_counter, void *stream) {
    if (!logits_bf16 || !act || !old_logp || !adv || !dlogits_bf16 || !loss_out || !db_head || B <= 0) return TA_E_INVALID;
    if (((uintptr_t)logits_bf16 | (uintptr_t)dlogits_bf16) & 15u) return TA_E_INVALID;
    ppo_actor_loss_kernel<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)logits_bf16, act, old_logp, adv, B, clip, ent_coef,
                                                                      (__nv_bfloat16 *)dlogits_bf16, loss_out, db_head, step_counter);
    return launch_ok("ppo_actor_loss_kernel");
}

int ta_ppo_critic_loss(const void *v_bf16, const float *target, int B, void *dv_bf16, float *loss_out, float *db_head,
                       float *step_counter, void *stream) {
    if (!v_bf16 || !target || !dv_bf16 || !loss_out || !db_head || B <= 0) return TA_E_INVALID;
    if (((uintptr_t)v_bf16 | (uintptr_t)dv_bf16) & 15u) return TA_E_INVALID;
    ppo_critic_loss_kernel<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)v_bf16, target, B, (__nv_bfloat16 *)dv_bf16, loss_out,
                                                                       db_head, step_counter);
    return launch_ok("ppo_critic_loss_kernel");
}

int ta_adam_shadow(float *p, const float *g, float *m, float *v, void *p_bf16, int64_t n, const float *step_counter, float lr,
                   double beta1, double beta2, float eps, float grad_scale, void *stream) {
    if (!p || !g || !m || !v || !p_bf16 || !step_counter || n <= 0) return TA_E_INVALID;
    unsigned nb = blocks_for(n, 256 * 4);
    if (nb > 148u * 8u) nb = 148u * 8u;
    adam_shadow_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16 *)p_bf16, n, step_counter, lr, beta1, beta2, eps, grad_scale);
    return launch_ok("adam_shadow_kernel");
}

int ta_tinet_prep(const ta_tinet_prep_args *args, void *stream) {
    if (!args || !args->w1 || !args->b1 || !args->w4 || !args->b4 || !args->fc0 || !args->fc0p || !args->pos || !args->pos16 || !args->head ||
        !args->head_b || !args->head8 || !args->head_b8 || args->nh < 1 || args->nh > 8)
        return TA_E_INVALID;
    PrepArgs a;
    memcpy(&a, args, sizeof(a));
    tinet_prep_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_ok("tinet_prep_kernel");
}

int ta_tinet_grad(const ta_tinet_grad_args *args, void *stream) {
    if (!args || args->nh < 1 || args->nh > 8) return TA_E_INVALID;
    if (args->dw4 && (!args->db4 || !args->g_w1 || !args->g_b1)) return TA_E_INVALID;
    if ((args->fc0p && !args->g_fc0) || (args->pos16 && !args->g_pos) || (args->head8 && !args->g_head)) return TA_E_INVALID;
    for (int k = 0; k < 4; k++)
        if (args->n[k] < 0 || (args->n[k] > 0 && (!args->src[k] || !args->dst[k]))) return TA_E_INVALID;
    GradArgs a;
    memcpy(&a, args, sizeof(a));
    tinet_grad_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_ok("tinet_grad_kernel");
}

int ta_gather_minibatch(const uint8_t *s, const float *p, const float *g, const int64_t *a, const float *old_logp, const float *adv,
                        const float *target_v, const int64_t *idx, const int64_t *src, int bs, uint8_t *sb, void *pg16_bf16, int32_t *a_mb,
                        float *old_mb, float *adv_mb, float *tv_mb, void *stream) {
    if (!s || !p || !g || !a || !old_logp || !adv || !target_v || !idx || bs <= 0 || !sb || !pg16_bf16 || !a_mb || !old_mb || !adv_mb || !tv_mb)
        return TA_E_INVALID;
    if ((uintptr_t)sb & 3u) return TA_E_INVALID;
    gather_minibatch_kernel<<<(unsigned)bs, 128, 0, (cudaStream_t)stream>>>(s, p, g, (const long long *)a, old_logp, adv, target_v,
                                                                          (const long long *)idx, (const long long *)src, bs, sb,
                                                                          (__nv_bfloat16 *)pg16_bf16, a_mb, old_mb, adv_mb, tv_mb);
    return launch_ok("gather_minibatch_kernel");
}

int ta_set_timing(ta_handle h, int on) {
    if (!h) return TA_E_INVALID;
    h->timing = on;
    return TA_OK;
}

int ta_last_step_ms(ta_handle h, float *ms) {
    if (!h || !ms) return TA_E_INVALID;
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return TA_OK;
}

/* measurement probe (not part of the ABI): what does this GPU sustain for a pure WRITE stream?
 * mode 0: 16-byte stores, one grid-stride pass; mode 1: 3 KB TMA bulk stores from shared memory
 * (the pattern of the step kernel's obs pass).  Used by scripts/write_bw.py. */
__global__ void __launch_bounds__(256) write_probe_stg(uint4 *dst, long long n16) {
    const uint4 v = make_uint4(0x01020304u, 0x05060708u, 0x090a0b0cu, 0x0d0e0f10u);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void __launch_bounds__(32) write_probe_bulk(uint8_t *dst, long long nchunks) {
    __shared__ __align__(128) uint8_t buf[3072];
    for (int i = threadIdx.x; i < 3072 / 4; i += 32) reinterpret_cast<uint32_t *>(buf)[i] = 0x01020304u * (i + 1);
    fence_proxy_async();
    __syncwarp();
    if (threadIdx.x == 0) {
        for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
            bulk_s2g(dst + c * 3072, buf, 3072);
            bulk_commit();
            bulk_wait_read<8>();
        }
        bulk_wait_read<0>();
    }
}
// mode 2 / 3: the store pattern of conv1_fwd_ws_kernel without its compute -- one CTA per SM, tile = 14 image rows of 4224
// bytes (two groups of 7: even / odd rows), one bulk store per row from a 1024-aligned staging block, at most `depth` tiles
// of stores outstanding per group (mode 2: depth 2 = the kernel's double buffer; mode 3: depth 8)
__global__ void __launch_bounds__(64) write_probe_rows(uint8_t *dst, long long ntiles, int depth) {
    extern __shared__ __align__(1024) uint8_t rows_buf[];
    for (int i = threadIdx.x; i < 7 * 5120 / 4; i += 64) reinterpret_cast<uint32_t *>(rows_buf)[i] = 0x01020304u * (i + 1);
    fence_proxy_async();
    __syncthreads();
    const int g = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            for (int r = 0; r < 7; r++) bulk_s2g(dst + ((t * 14 + 2 * r + g) * 4224), rows_buf + r * 5120, 4224);
            bulk_commit();
            if (depth <= 2) bulk_wait_read<1>();
            else bulk_wait_read<7>();
        }
        bulk_wait_read<0>();
    }
}
int ta_debug_write_probe(void *dst, int64_t bytes, int mode, void *stream) {
    if (!dst || bytes <= 0) return TA_E_INVALID;
    if (mode >= 2) {
        CK(cudaFuncSetAttribute(write_probe_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 5120));
        write_probe_rows<<<148, 64, 7 * 5120, (cudaStream_t)stream>>>((uint8_t *)dst, bytes / (14 * 4224), mode == 2 ? 2 : 8);
        return launch_ok("write_probe_rows");
    }
    if (mode == 0) write_probe_stg<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((uint4 *)dst, bytes / 16);
    else write_probe_bulk<<<148 * 16, 32, 0, (cudaStream_t)stream>>>((uint8_t *)dst, bytes / 3072);
    return launch_ok("write_probe");
}

/* test hook: route ta_conv1_fwd through the tcgen05 kernel (1, default) or the FP32-FMA kernel (0); returns the
 * previous setting (-1 = not decided yet) */
int ta_debug_conv1_tc(int on) {   // 0: FP32-FMA kernel, 1: conv1_fwd_tc_kernel, 2: conv1_fwd_ws_kernel (default); -1: back to TA_CONV1_TC / default
    const int prev = g_use_tc;
    g_use_tc = on < 0 ? -1 : (on > 2 ? 2 : on);
    return prev;
}

/* test hook for the tcgen05 weight-gradient kernel: on/off (returns the previous setting) */
int ta_debug_push_tma(int on) {   // 0: register kernel (stack_push_codes_tile_kernel), 1: stack_push_tma_kernel; -1: TA_PUSH_TMA / default
    g_push_tma = on < 0 ? -1 : (on != 0);
    return TA_OK;
}

int ta_debug_conv1_bwd_tc(int on) {
    const int prev = g_bwd_tc;
    g_bwd_tc = on != 0;
    return prev;
}

/* check for the tcgen05 conv1 kernel: 1 if any launch gave up waiting for its MMA (synchronises) */
int ta_debug_conv1_tc_failed(void) {
    int *g_tc_fail = g_tc_fail_dev[cur_dev()];   // (the current device's flag)
    if (!g_tc_fail) return 0;
    int v = 0;
    if (cudaMemcpy(&v, g_tc_fail, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (v != 0) cudaMemset(g_tc_fail, 0, sizeof(int));  // reported once: later updates start clean
    return v;
}

/* test hook: emit the observations with per-lane stores (1) instead of TMA bulk stores
 * (0, default) -- the path taken when an obs slice is not 16-byte aligned */
int ta_debug_force_generic_obs(int on) {
    g_force_generic = on;
    return TA_OK;
}

}  // extern "C"
