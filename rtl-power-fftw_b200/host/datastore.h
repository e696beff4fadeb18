// datastore.h -- Datastore with the reference's public surface
// (src/datastore.h:35-68): the same members the producer touches -- pwr,
// repeats_done, acquisition_finished, status_mutex, status_change,
// empty_buffers, occupied_buffers, queue_histogram, fftThread(),
// printQueueHistogram() -- but the FFTW plan/arrays are replaced by an rpf_ctx
// (include/rpf_b200.h) and the worker body by rpf_push().
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rpf_b200.h"
#include "rpf_params.h"

using Buffer = std::vector<uint8_t>;

class Datastore {
  public:
    const Params &params;
    int64_t repeats_done = 0;

    std::mutex status_mutex;
    // guarded by status_mutex (same convention as the reference, src/datastore.h:41-42)
    std::deque<Buffer *> empty_buffers;
    std::deque<Buffer *> occupied_buffers;
    bool acquisition_finished = false;
    std::condition_variable status_change;
    std::vector<int> queue_histogram;

    std::vector<float> &window_values;
    std::vector<double> pwr;

    Datastore(const Params &params, std::vector<float> &window_values);
    ~Datastore();
    Datastore(const Datastore &) = delete;
    Datastore &operator=(const Datastore &) = delete;

    // Runs in its own thread, one per acquisition (src/acquisition.cxx:256).
    // Never throws: a CUDA failure is latched and the thread keeps recycling
    // buffers so the producer cannot deadlock; call rethrow_worker_error()
    // after join().
    void fftThread();
    void printQueueHistogram() const;

    // Throws RPFexception(..., ReturnValue::HardwareError) if the last
    // fftThread() run hit an error.  Call on the producer thread after join().
    void rethrow_worker_error() const;

    rpf_ctx *context() const { return ctx; }

  private:
    rpf_ctx *ctx = nullptr;
    std::vector<Buffer *> all_buffers;
    int worker_status = 0;
    std::string worker_error;
};
