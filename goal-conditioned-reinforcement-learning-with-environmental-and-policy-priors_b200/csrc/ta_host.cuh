// ta_host.cuh -- host side of ta_step_host: the decode stage that turns the packed transfer form of the
// observations (2-bit cell codes, 16 cells per word, written by step_obs_kernel with flags bit 5) into the caller's
// uint8 [n][V][V][3] array (Grid.encode's bytes, gym_minigrid/minigrid.py:749-772, what env.step returns as
// obs["image"], minigrid.py:1439-1441), and the one-byte step status into reward / terminated / truncated.
//
// Why: the reference-facing call hands HOST arrays back.  Shipping the expanded observation over PCIe moves 867 B per
// env-step (57 MB per 65536-env step at ~54 GB/s: 1.05 ms); the packed form is 72.25 B per env-step (4.7 MB), and the
// expansion is a byte-LUT pass that a few host threads run at memory-write speed while the next chunk is still
// in flight.  This is a stage of the host call, not a fallback: the env transition and gen_obs always run on the GPU.
//
// Host C++ only (compiled by nvcc's host compiler); SSE2 is part of the x86-64 baseline.
#pragma once
#include <emmintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace ta_host {

// LUT[b] = the 12 obs bytes of the 4 cells packed in byte b (cell q = bits 2q, 2q+1), padded to 16
struct alignas(16) Lut {
    uint8_t e[256][16];
    Lut() {
        static const uint8_t type[4] = {1, 2, 6, 8}, color[4] = {0, 5, 4, 1};  // empty, wall, ball (yellow), goal (green)
        for (int b = 0; b < 256; b++) {
            memset(e[b], 0, 16);
            for (int q = 0; q < 4; q++) {
                const int c = (b >> (2 * q)) & 3;
                e[b][3 * q] = type[c];
                e[b][3 * q + 1] = color[c];
            }
        }
    }
};
inline const Lut &lut() {
    static const Lut L;
    return L;
}

// 16 cells (one code word) -> 48 bytes as three 16-byte vectors
static inline void expand_word(uint32_t w, const Lut &L, __m128i &o0, __m128i &o1, __m128i &o2) {
    const __m128i a = _mm_load_si128(reinterpret_cast<const __m128i *>(L.e[w & 0xFFu]));
    const __m128i b = _mm_load_si128(reinterpret_cast<const __m128i *>(L.e[(w >> 8) & 0xFFu]));
    const __m128i c = _mm_load_si128(reinterpret_cast<const __m128i *>(L.e[(w >> 16) & 0xFFu]));
    const __m128i d = _mm_load_si128(reinterpret_cast<const __m128i *>(L.e[w >> 24]));
    o0 = _mm_or_si128(a, _mm_slli_si128(b, 12));
    o1 = _mm_or_si128(_mm_srli_si128(b, 4), _mm_slli_si128(c, 8));
    o2 = _mm_or_si128(_mm_srli_si128(c, 8), _mm_slli_si128(d, 4));
}

// words[0..nwords) -> dst[0..48*nwords), of which only the first `valid` bytes exist (ragged last tile)
static inline void expand_run(const uint32_t *words, long long nwords, uint8_t *dst, long long valid) {
    const Lut &L = lut();
    long long full = valid / 48;
    if (full > nwords) full = nwords;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {  // streaming stores: the 57 MB result is not read back by this core
        for (long long r = 0; r < full; r++) {
            __m128i o0, o1, o2;
            expand_word(words[r], L, o0, o1, o2);
            __m128i *d = reinterpret_cast<__m128i *>(dst + 48 * r);
            _mm_stream_si128(d, o0);
            _mm_stream_si128(d + 1, o1);
            _mm_stream_si128(d + 2, o2);
        }
    } else {
        for (long long r = 0; r < full; r++) {
            __m128i o0, o1, o2;
            expand_word(words[r], L, o0, o1, o2);
            __m128i *d = reinterpret_cast<__m128i *>(dst + 48 * r);
            _mm_storeu_si128(d, o0);
            _mm_storeu_si128(d + 1, o1);
            _mm_storeu_si128(d + 2, o2);
        }
    }
    if (full < nwords && valid > 48 * full) {  // the partial word at the end of a ragged tile
        alignas(16) uint8_t tmp[48];
        __m128i o0, o1, o2;
        expand_word(words[full], L, o0, o1, o2);
        _mm_store_si128(reinterpret_cast<__m128i *>(tmp), o0);
        _mm_store_si128(reinterpret_cast<__m128i *>(tmp + 16), o1);
        _mm_store_si128(reinterpret_cast<__m128i *>(tmp + 32), o2);
        memcpy(dst + 48 * full, tmp, (size_t)(valid - 48 * full));
    }
}

static const uint32_t REWARD_BITS[8] = {0xBC23D70Au, 0xBDCCCCCDu, 0xBF666666u, 0x3E4CCCCDu, 0x3F666666u, 0, 0, 0};  // ta_common.cuh reward_value

struct Job {
    const uint32_t *codes = nullptr;  // [ntiles][runs]
    const uint8_t *status = nullptr;  // [ntiles * 32]
    uint8_t *obs = nullptr;
    float *reward = nullptr;
    uint8_t *term = nullptr, *trunc = nullptr;
    long long n = 0, ntiles = 0, nunits = 0;
    int runs = 0, obs_bytes = 0, unit_tiles = 1;
    std::atomic<long long> next{0}, ready{0}, finished{0};  // in units
};

inline void decode_unit(const Job &j, long long u) {
    const long long t0 = u * j.unit_tiles, t1 = t0 + j.unit_tiles < j.ntiles ? t0 + j.unit_tiles : j.ntiles;
    for (long long t = t0; t < t1; t++) {
        long long nvalid = j.n - t * 32;
        if (nvalid > 32) nvalid = 32;
        expand_run(j.codes + t * j.runs, j.runs, j.obs + t * 32 * (long long)j.obs_bytes, nvalid * j.obs_bytes);
        for (long long e = t * 32; e < t * 32 + nvalid; e++) {
            const uint8_t s = j.status[e];
            memcpy(j.reward + e, &REWARD_BITS[s & 7u], 4);
            j.term[e] = (s >> 3) & 1u;
            j.trunc[e] = (s >> 4) & 1u;
        }
    }
}

inline void work(Job &j) {
    for (;;) {
        const long long u = j.next.fetch_add(1, std::memory_order_relaxed);
        if (u >= j.nunits) break;
        int spins = 0;
        while (u >= j.ready.load(std::memory_order_acquire)) {  // this unit's chunk is still on the wire
            _mm_pause();
            if (++spins > 4096) { std::this_thread::yield(); spins = 0; }
        }
        decode_unit(j, u);
        j.finished.fetch_add(1, std::memory_order_release);
    }
}

// A small persistent pool: workers spin briefly for the next job (back-to-back steps), then sleep.
class Pool {
  public:
    explicit Pool(int nthreads) : stop_(false), epoch_(0), job_(nullptr), active_(0) {
        for (int i = 0; i < nthreads; i++) threads_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            epoch_.fetch_add(1);
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    int size() const { return (int)threads_.size(); }
    void start(Job *j) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = j;
            active_.store((int)threads_.size(), std::memory_order_relaxed);
            epoch_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
    }
    void finish(Job *j) {  // the caller works too, then waits for the stragglers
        work(*j);
        while (j->finished.load(std::memory_order_acquire) < j->nunits || active_.load(std::memory_order_acquire) > 0) _mm_pause();
    }

  private:
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            int spins = 0;
            while (epoch_.load(std::memory_order_acquire) == seen) {
                _mm_pause();
                if (++spins > 200000) {  // ~1 ms of polling, then block
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
                    break;
                }
            }
            seen = epoch_.load(std::memory_order_acquire);
            if (stop_) return;
            Job *j = job_;
            if (j) work(*j);
            active_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_;
    std::atomic<unsigned long long> epoch_;
    Job *job_;
    std::atomic<int> active_;
};

inline int default_threads() {
    if (const char *e = getenv("TA_HOST_THREADS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 256) return v;
    }
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    int cores = (int)std::thread::hardware_concurrency();
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) cores = CPU_COUNT(&allowed);
    int share = 1;  // one process per GPU: split the host cores between the ranks of this node
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = atoi(e) > 0 ? atoi(e) : 1;
    int t = cores / share;
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    return t;
}

}  // namespace ta_host
