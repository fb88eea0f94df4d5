// h2d_staged.h — host-to-device copy of a big PAGEABLE array at more than one core's memcpy rate.
//
// cudaMemcpyAsync from pageable memory stages through the driver's own bounce buffer on the calling thread: ~8 GB/s, so
// the 1.5 GB of a 10 M-triangle scene's staging arrays (vertices, elements, per-primitive metadata) took 190 ms of a
// 290 ms commit.  Here a few host threads copy interleaved chunks into pinned buffers of the library (allocated once per
// process and device, 96 MB) and queue one DMA per chunk on the caller's stream, so the host copy runs on several cores and
// overlaps the DMA of the previous chunks.  Semantics = the pageable cudaMemcpyAsync it replaces: on return the source has
// been read completely (the caller may free it), the copy itself is ordered on `stream`.
// CRB_STAGED_H2D=0 falls back to the plain call (A/B switch).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace crb {

class StagedCopier {
public:
    static constexpr size_t CHUNK = (size_t)8 << 20;
    static constexpr int THREADS = 6, DEPTH = 2;
    static constexpr size_t MIN_BYTES = (size_t)32 << 20;  // below this the plain call wins

    static cudaError_t copy(int device, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
        static const bool enabled = [] {
            const char* e = std::getenv("CRB_STAGED_H2D");
            return !(e && e[0] == '0');
        }();
        if (!enabled || bytes < MIN_BYTES || device < 0 || device >= 64) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
        StagedCopier& c = slot(device);
        std::lock_guard<std::mutex> lk(c.mu_);  // one staged copy per device at a time (the buffers are shared)
        if (!c.ready_ && !c.init(device)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
        const size_t n_chunks = (bytes + CHUNK - 1) / CHUNK;
        cudaError_t errs[THREADS];
        auto work = [&](int t) {
            errs[t] = cudaSetDevice(device);
            int b = 0;
            for (size_t k = (size_t)t; k < n_chunks && errs[t] == cudaSuccess; k += THREADS, b = (b + 1) % DEPTH) {
                const size_t off = k * CHUNK, n = bytes - off < CHUNK ? bytes - off : CHUNK;
                char* buf = c.buf_[t][b];
                if ((errs[t] = cudaEventSynchronize(c.ev_[t][b])) != cudaSuccess) break;  // the DMA that last read this buffer is done
                std::memcpy(buf, static_cast<const char*>(src) + off, n);
                if ((errs[t] = cudaMemcpyAsync(static_cast<char*>(dst) + off, buf, n, cudaMemcpyHostToDevice, stream)) != cudaSuccess) break;
                errs[t] = cudaEventRecord(c.ev_[t][b], stream);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < THREADS; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        for (int t = 0; t < THREADS; ++t)
            if (errs[t] != cudaSuccess) return errs[t];
        return cudaSuccess;
    }

private:
    static StagedCopier& slot(int device) {
        static StagedCopier* s = new StagedCopier[64];  // never destroyed: pinned memory must not be freed after the CUDA runtime
        return s[device];
    }
    bool init(int device) {
        if (cudaSetDevice(device) != cudaSuccess) return false;
        char* base = nullptr;
        if (cudaHostAlloc(reinterpret_cast<void**>(&base), CHUNK * THREADS * DEPTH, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        for (int t = 0; t < THREADS; ++t)
            for (int b = 0; b < DEPTH; ++b) {
                buf_[t][b] = base + CHUNK * (size_t)(t * DEPTH + b);
                if (cudaEventCreateWithFlags(&ev_[t][b], cudaEventDisableTiming) != cudaSuccess) return false;
            }
        ready_ = true;
        return true;
    }
    std::mutex mu_;
    bool ready_ = false;
    char* buf_[THREADS][DEPTH] = {};
    cudaEvent_t ev_[THREADS][DEPTH] = {};
};

}  // namespace crb
