/*
 * ref_driver.cpp -- drives the reference's own Datastore (compiled unmodified
 * from /root/reference/src/datastore.cxx) from a memory / generator byte
 * source.  TEST INFRASTRUCTURE ONLY (oracle/README.md).
 *
 * What is the reference's own code here: class Datastore, its constructor,
 * destructor and fftThread() -- the whole hot path (src/datastore.cxx:23-96).
 * What this file supplies in the reference's place:
 *   - Params::Params(int, char**): the only declared constructor lives in
 *     src/params.cxx, which needs TCLAP (absent).  Defined empty here; fields
 *     keep their in-class defaults (src/params.h:33-66) and are then set.
 *   - the producer loop of Acquisition::run() (src/acquisition.cxx:252-256,
 *     271-347) restated, reading from a callback instead of Rtlsdr::read()
 *     (src/device.cxx:92-97; librtlsdr absent).
 *   - the FFT behind <fftw3.h>: shim_fft.cpp (stand-in, not FFTW).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "datastore.h"   // /root/reference/src/datastore.h via -I

Params::Params(int, char **) {}

extern "C" {

typedef void (*rpf_ref_read_fn)(uint8_t *dst, size_t len, uint64_t stream_offset, void *user);

struct mem_source { const uint8_t *bytes; size_t len; int wrap; };

static void mem_read(uint8_t *dst, size_t len, uint64_t off, void *user)
{
    const mem_source *m = (const mem_source *)user;
    if (!m->wrap) {
        size_t avail = off < m->len ? m->len - (size_t)off : 0;
        size_t n = std::min(avail, len);
        if (n) memcpy(dst, m->bytes + off, n);
        if (n < len) memset(dst + n, 127, len - n);   // zero-signal padding
        return;
    }
    size_t pos = (size_t)(off % m->len), done = 0;
    while (done < len) {
        size_t n = std::min(len - done, m->len - pos);
        memcpy(dst + done, m->bytes + pos, n);
        done += n; pos = 0;
    }
}

/* One Acquisition::run() worth of work.  Returns 0, fills pwr_out[N],
 * *repeats_done, hist_out[buffers+1] (may be NULL), *reads (may be NULL). */
int rpf_ref_run_cb(int N, int64_t repeats, const float *window, int buffers, int buf_length,
                   rpf_ref_read_fn read_fn, void *user,
                   double *pwr_out, int64_t *repeats_done, int *hist_out, int64_t *reads)
{
    if (N <= 0 || (N & 1) || repeats < 0 || buffers < 1 || buf_length < base_buf) return -1;
    Params params(0, nullptr);
    params.N = N;
    params.repeats = repeats;
    params.buffers = buffers;
    params.buf_length = buf_length;
    params.window = window != nullptr;
    std::vector<float> window_values;
    if (window) window_values.assign(window, window + N);

    Datastore data(params, window_values);          // src/rtl_power_fftw.cxx:112

    // --- Acquisition::run(), src/acquisition.cxx:252-256
    std::fill(data.pwr.begin(), data.pwr.end(), 0);
    data.acquisition_finished = false;
    data.repeats_done = 0;
    std::thread t(&Datastore::fftThread, &data);

    // --- producer loop, src/acquisition.cxx:271-332
    std::unique_lock<std::mutex> status_lock(data.status_mutex, std::defer_lock);
    int64_t dataTotal = 2 * (int64_t)params.N * params.repeats;
    int64_t dataRead = 0, nreads = 0;
    while (dataRead < dataTotal) {
        status_lock.lock();
        data.queue_histogram[data.empty_buffers.size()]++;
        while (data.empty_buffers.empty())
            data.status_change.wait(status_lock);
        Buffer &buffer(*data.empty_buffers.front());
        data.empty_buffers.pop_front();
        status_lock.unlock();

        int64_t dataNeeded = dataTotal - dataRead;
        if (dataNeeded >= params.buf_length)
            dataNeeded = params.buf_length;
        else {
            dataNeeded = base_buf * (int64_t)ceil((double)dataNeeded / base_buf);
            if (dataNeeded > params.buf_length)
                dataNeeded = params.buf_length;
        }
        buffer.resize(dataNeeded);
        read_fn(buffer.data(), buffer.size(), (uint64_t)dataRead, user);   // Rtlsdr::read stand-in
        nreads++;
        dataRead += dataNeeded;
        status_lock.lock();
        data.occupied_buffers.push_back(&buffer);
        data.status_change.notify_all();
        status_lock.unlock();
    }
    // --- src/acquisition.cxx:343-347
    status_lock.lock();
    data.acquisition_finished = true;
    data.status_change.notify_all();
    status_lock.unlock();
    t.join();

    memcpy(pwr_out, data.pwr.data(), sizeof(double) * N);
    *repeats_done = data.repeats_done;
    if (hist_out) for (int i = 0; i <= buffers; i++) hist_out[i] = data.queue_histogram[i];
    if (reads) *reads = nreads;
    return 0;
}

/* Memory-backed source.  wrap != 0: the stream repeats with period
 * stream_len; wrap == 0: bytes past the end read as 127 (zero signal). */
int rpf_ref_run(int N, int64_t repeats, const float *window, int buffers, int buf_length,
                const uint8_t *stream, size_t stream_len, int wrap,
                double *pwr_out, int64_t *repeats_done, int *hist_out, int64_t *reads)
{
    if (!stream || stream_len == 0) return -1;
    mem_source m{stream, stream_len, wrap};
    return rpf_ref_run_cb(N, repeats, window, buffers, buf_length, mem_read, &m,
                          pwr_out, repeats_done, hist_out, reads);
}

/* `workers` independent reference instances (each its own Datastore, producer
 * and worker thread) over the same periodic stream: the "all host cores"
 * baseline.  pwr_out receives instance 0's result. */
int rpf_ref_run_parallel(int workers, int N, int64_t repeats_each, const float *window, int buffers,
                         int buf_length, const uint8_t *stream, size_t stream_len,
                         double *pwr_out, int64_t *repeats_done_total)
{
    if (workers < 1) return -1;
    std::vector<std::thread> th;
    std::vector<std::vector<double>> pw(workers, std::vector<double>(N));
    std::vector<int64_t> done(workers, 0);
    std::vector<int> rc(workers, 0);
    for (int w = 0; w < workers; w++)
        th.emplace_back([&, w] {
            rc[w] = rpf_ref_run(N, repeats_each, window, buffers, buf_length, stream, stream_len, 1,
                                pw[w].data(), &done[w], nullptr, nullptr);
        });
    for (auto &t : th) t.join();
    int64_t tot = 0;
    for (int w = 0; w < workers; w++) { if (rc[w]) return rc[w]; tot += done[w]; }
    memcpy(pwr_out, pw[0].data(), sizeof(double) * N);
    *repeats_done_total = tot;
    return 0;
}

}  // extern "C"
