#include "datastore.h"

#include <iostream>

Datastore::Datastore(const Params &p, std::vector<float> &win)
    : params(p), queue_histogram(p.buffers + 1, 0), window_values(win), pwr(p.N)
{
    // plan creation: was fftwf_alloc_complex x2 + fftwf_plan_dft_1d (src/datastore.cxx:30-33)
    rpf_config cfg = {};
    cfg.N = params.N;
    cfg.device = params.cuda_device;
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

        lock.lock();
        empty_buffers.push_back(buffer);
        status_change.notify_all();
        lock.unlock();
    }
    if (worker_status == 0) {
        rc = rpf_finish(ctx, pwr.data(), &repeats_done);
        if (rc != RPF_OK) { worker_status = rc; worker_error = rpf_strerror(ctx); }
    }
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
