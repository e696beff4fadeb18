#include "datastore.h"

#include <iostream>

Datastore::Datastore(const Params &p, std::vector<float> &win)
    : params(p), queue_histogram(p.buffers + 1, 0), window_values(win), pwr(p.N)
{
    // plan creation: was fftwf_alloc_complex x2 + fftwf_plan_dft_1d (src/datastore.cxx:30-33)
    rpf_config cfg = {};
    cfg.N = params.N;
    cfg.device = params.cuda_device;
    if (params.cuda_devices.size() > RPF_MAX_DEVICES)
        throw RPFexception("too many CUDA devices", ReturnValue::InvalidArgument);
    cfg.n_devices = (int32_t)params.cuda_devices.size();   // 0: the single GPU above
    for (size_t i = 0; i < params.cuda_devices.size(); i++) cfg.devices[i] = params.cuda_devices[i];
    if (params.pipelined_push) cfg.flags |= RPF_FLAG_PIPELINED_PUSH;
    cfg.window = params.window ? window_values.data() : nullptr;
    cfg.buf_length = params.buf_length;
    if (params.window && (int)window_values.size() != params.N)
        throw RPFexception("window function must have N values", ReturnValue::InvalidInput);
    int rc = rpf_create(&cfg, &ctx);
    if (rc != RPF_OK)
        throw RPFexception(std::string("FFT worker: ") + rpf_strerror(nullptr),
                           rc == RPF_ERR_INVALID_ARG || rc == RPF_ERR_UNSUPPORTED_N ? ReturnValue::InvalidArgument
                                                                                    : ReturnValue::HardwareError);
    // buffer pool (src/datastore.cxx:27-28); page-locked so reads land in DMA-able memory
    for (int i = 0; i < params.buffers; i++) {
        Buffer *b = new Buffer(params.buf_length);
        rpf_host_register(b->data(), b->capacity());   // best effort: pageable still works
        all_buffers.push_back(b);
        empty_buffers.push_back(b);
    }
}

Datastore::~Datastore()
{
    rpf_destroy(ctx);                                   // fftwf_destroy_plan + fftwf_free (src/datastore.cxx:43-45)
    for (Buffer *b : all_buffers) {
        rpf_host_unregister(b->data());
        delete b;
    }
}

void Datastore::fftThread()
{
    worker_status = 0;
    worker_error.clear();
    // the state a fresh worker starts from: pwr = 0, repeats_done = 0 (src/acquisition.cxx:252-254),
    // fft_pointer = 0 (src/datastore.cxx:52)
    int rc = rpf_reset(ctx, params.repeats);
    if (rc != RPF_OK) { worker_status = rc; worker_error = rpf_strerror(ctx); }

    // RPF_FLAG_PIPELINED_PUSH: the copy of buffer k is still in flight when rpf_push returns, so buffer k goes
    // back to the empty queue only after the push of buffer k+1 (or the final rpf_finish) has returned.
    Buffer *in_flight = nullptr;
    auto recycle = [this](Buffer *b) {
        std::lock_guard<std::mutex> g(status_mutex);
        empty_buffers.push_back(b);
        status_change.notify_all();
    };
    std::unique_lock<std::mutex> lock(status_mutex, std::defer_lock);
    for (;;) {
        lock.lock();
        status_change.wait(lock, [this] { return !occupied_buffers.empty() || acquisition_finished; });
        if (occupied_buffers.empty()) { lock.unlock(); break; }   // finished and drained
        Buffer *buffer = occupied_buffers.front();
        occupied_buffers.pop_front();
        lock.unlock();

        if (worker_status == 0) {
            // the whole of src/datastore.cxx:66-89 for this buffer
            rc = rpf_push(ctx, buffer->data(), buffer->size());
            if (rc != RPF_OK) { worker_status = rc; worker_error = rpf_strerror(ctx); }
        }
        // on error: keep returning buffers so Acquisition::run() never blocks forever

        if (params.pipelined_push && worker_status == 0) {
            if (in_flight) recycle(in_flight);
            in_flight = buffer;
            // with a single buffer in the pool the producer would starve: release it at once (after a sync)
            if (params.buffers < 2) { rpf_sync(ctx); recycle(in_flight); in_flight = nullptr; }
        } else {
            if (in_flight) { rpf_sync(ctx); recycle(in_flight); in_flight = nullptr; }
            recycle(buffer);
        }
    }
    if (worker_status == 0) {
        rc = rpf_finish(ctx, pwr.data(), &repeats_done);
        if (rc != RPF_OK) { worker_status = rc; worker_error = rpf_strerror(ctx); }
    } else {
        rpf_sync(ctx);
    }
    if (in_flight) recycle(in_flight);
}

void Datastore::rethrow_worker_error() const
{
    if (worker_status != 0)
        throw RPFexception("FFT worker failed: " + worker_error, ReturnValue::HardwareError);
}

void Datastore::printQueueHistogram() const
{
    std::cerr << "Buffer queue histogram: ";
    for (int n : queue_histogram) std::cerr << n << " ";
    std::cerr << std::endl;
}
