#include "acquisition.h"

#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <iostream>
#include <thread>

// Text grammar of src/acquisition.cxx:32-58, including its corner case: the value kept for a line is whatever the
// extraction loop leaves in `value` once it stops -- and a FAILED extraction stores 0 (C++11 num_get), so a line such
// as "0.5 abc" yields 0, not 0.5, while "1 2 3" yields 3.  (Pinned against the reference's own function compiled from
// /root/reference: tests/test_aux_parser.py.)
template <typename T> std::vector<T> read_inputfile(std::istream &stream)
{
    std::vector<T> values;
    std::string line;
    while (std::getline(stream, line)) {
        std::istringstream ls(line);
        if ((ls >> std::ws).peek() == '#') continue;          // comment line
        T value = T();
        unsigned parsed = 0;
        while (ls >> value) parsed++;                         // the failing extraction that ends the loop zeroes `value`
        if (parsed > 0) values.push_back(value);              // lines without a number are skipped
    }
    return values;
}
template std::vector<float> read_inputfile<float>(std::istream &);
template std::vector<double> read_inputfile<double>(std::istream &);

AuxData::AuxData(const Params &params) { load(params, std::cin); }
AuxData::AuxData(const Params &params, std::istream &in) { load(params, in); }

void AuxData::load(const Params &params, std::istream &in)
{
    const std::string N = std::to_string(params.N);
    if (params.window && params.baseline && params.window_file == "-" && params.baseline_file == "-") {
        // Both from stdin: 2N values, window first, then baseline (the code's stated intent,
        // src/acquisition.cxx:64-69).  NOTE: the reference computes the split point from the still
        // empty window vector (src/acquisition.cxx:76), so there every value lands in the baseline and
        // the window stays empty; that defect is deliberately not reproduced.
        std::vector<double> values = read_inputfile<double>(in);
        if ((int)values.size() != 2 * params.N)
            throw RPFexception("Error reading window function and baseline from stdin. Expected " +
                               std::to_string(2 * params.N) + " values, found " + std::to_string(values.size()) + ".",
                               ReturnValue::InvalidInput);
        for (int i = 0; i < params.N; i++) window_values.push_back((float)values[i]);
        for (int i = params.N; i < 2 * params.N; i++) baseline_values.push_back(values[i]);
        return;
    }
    // ONE ifstream serves both files, as in the reference (src/acquisition.cxx:62,108,132): after the window file
    // has been read to its end, open() on the still-open stream fails, so "window from a file AND baseline from a
    // file" ends in "Could not open <baseline_file>" there -- at most one of the two can come from a file, the
    // other from stdin.  Kept as is (same error text and exit code); see DESIGN.md section 8.
    std::ifstream fs;
    auto open_and_read = [&](const std::string &file, bool as_float) {
        std::istream *stream = &in;
        if (file != "-") {
            fs.open(file);
            if (!fs.good()) throw RPFexception("Could not open " + file + ". Quitting.", ReturnValue::InvalidInput);
            stream = &fs;
        }
        if (as_float) window_values = read_inputfile<float>(*stream);
        else baseline_values = read_inputfile<double>(*stream);
    };
    if (params.window) {
        open_and_read(params.window_file, true);
        if ((int)window_values.size() != params.N)
            throw RPFexception("Error reading window function. Expected " + N + " values, found " +
                               std::to_string(window_values.size()) + ".", ReturnValue::InvalidInput);
    }
    if (params.baseline) {
        open_and_read(params.baseline_file, false);
        if ((int)baseline_values.size() != params.N)
            throw RPFexception("Error reading baseline. Expected " + N + " values, found " +
                               std::to_string(baseline_values.size()) + ".", ReturnValue::InvalidInput);
    }
}

bool SyntheticSource::read(Buffer &buffer)
{
    const uint64_t sd = per_frequency ? seed ^ ((uint64_t)freq * 0x9E3779B97F4A7C15ull) : seed;
    rpf_synth_fill_host(buffer.data(), buffer.size(), offset, sd, mode);
    offset += buffer.size();
    return true;
}

Plan::Plan(Params &p, int rate) : actual_samplerate(rate), params(p)
{
    // repeats from the integration time at the TRUE sample rate (src/acquisition.cxx:162-163)
    if (params.integration_time_isSet)
        params.repeats = (int64_t)std::ceil(actual_samplerate * params.integration_time / params.N);

    // buffer auto-sizing (src/acquisition.cxx:166-176): up to ~1.6 MB of data -> the smallest multiple of the USB
    // quantum that holds it, otherwise the default 100 quanta stay
    if (!params.buf_length_isSet) {
        int64_t mult = (int64_t)std::ceil((2.0 * params.N * params.repeats) / base_buf);
        if (mult <= default_buf_multiplier) params.buf_length = base_buf * (mult == 0 ? 1 : mult);
    }

    // centre frequencies (src/acquisition.cxx:180-198): the hops cover [startfreq, stopfreq] with at least
    // min_overlap percent of the bandwidth shared by neighbours; the surplus is spread evenly as extra overlap
    if (params.freq_hopping_isSet) {
        const double span = (double)(params.stopfreq - params.startfreq);
        const double min_overhang = actual_samplerate * params.min_overlap / 100;
        const int hops = (int)std::ceil((span - min_overhang) / ((double)actual_samplerate - min_overhang));
        if (hops > 1) {
            const int overhang = (int)(((int64_t)hops * actual_samplerate - (params.stopfreq - params.startfreq)) / (hops - 1));
            freqs_to_tune.push_back((int64_t)(params.startfreq + actual_samplerate / 2.0));
            for (int hop = 1; hop < hops; hop++) freqs_to_tune.push_back(freqs_to_tune.back() + actual_samplerate - overhang);
        } else {
            freqs_to_tune.push_back((params.startfreq + params.stopfreq) / 2);
        }
    } else {
        freqs_to_tune.push_back(params.cfreq);
    }
}

void Plan::print(std::ostream &err) const
{
    err << "Number of bins: " << params.N << std::endl;
    err << "Total number of (complex) samples to collect: " << (int64_t)params.N * params.repeats << std::endl;
    err << "Buffer length: " << params.buf_length << std::endl;
    err << "Number of averaged spectra: " << params.repeats << std::endl;
    err << "Estimated time of measurements: " << (double)params.N * params.repeats / actual_samplerate << " seconds"
        << std::endl;
}

Acquisition::Acquisition(const Params &p, AuxData &a, SampleSource &s, Datastore &d, int rate, int64_t f)
    : params(p), aux(a), source(s), data(d), actual_samplerate(rate), freq(f) {}

void Acquisition::run()
{
    source.set_frequency(freq);
    tuned_freq = source.frequency();

    // src/acquisition.cxx:252-256: reset, then one worker thread per acquisition
    std::fill(data.pwr.begin(), data.pwr.end(), 0);
    data.acquisition_finished = false;
    data.repeats_done = 0;
    std::thread worker(&Datastore::fftThread, &data);

    // producer loop, src/acquisition.cxx:271-332
    std::unique_lock<std::mutex> lock(data.status_mutex, std::defer_lock);
    const int64_t dataTotal = 2 * (int64_t)params.N * params.repeats;
    int64_t dataRead = 0;
    while (dataRead < dataTotal) {
        lock.lock();
        data.queue_histogram[data.empty_buffers.size()]++;       // back-pressure gauge
        data.status_change.wait(lock, [this] { return !data.empty_buffers.empty(); });
        Buffer &buffer = *data.empty_buffers.front();
        data.empty_buffers.pop_front();
        lock.unlock();

        int64_t dataNeeded = dataTotal - dataRead;
        if (dataNeeded >= params.buf_length) {
            dataNeeded = params.buf_length;
        } else {   // last read: round up to the USB quantum, capped at one buffer
            dataNeeded = base_buf * (int64_t)std::ceil((double)dataNeeded / base_buf);
            if (dataNeeded > params.buf_length) dataNeeded = params.buf_length;
        }
        buffer.resize(dataNeeded);

        const bool ok = source.read(buffer);
        deviceReadouts++;
        lock.lock();
        if (!ok) {
            std::cerr << "Error: dropped samples." << std::endl;
            data.empty_buffers.push_front(&buffer);               // retry with the same buffer
        } else {
            successfulReadouts++;
            dataRead += dataNeeded;
            data.occupied_buffers.push_back(&buffer);
            data.status_change.notify_all();
        }
        lock.unlock();
    }
    lock.lock();
    data.acquisition_finished = true;
    data.status_change.notify_all();
    lock.unlock();
    worker.join();
    data.rethrow_worker_error();   // CUDA failure -> RPFexception(HardwareError) on this thread
}

void Acquisition::print_summary() const { print_summary(std::cerr); }

void Acquisition::print_summary(std::ostream &err) const
{
    err << "Actual number of (complex) samples collected: " << (int64_t)params.N * data.repeats_done << std::endl;
    err << "Actual number of device readouts: " << deviceReadouts << std::endl;
    err << "Number of successful readouts: " << successfulReadouts << std::endl;
    err << "Actual number of averaged spectra: " << data.repeats_done << std::endl;
    err << "Effective integration time: " << (double)params.N * data.repeats_done / actual_samplerate << " seconds"
        << std::endl;
}

// Text output of src/acquisition.cxx:360-433 (non-matrix mode, without the
// timestamp header lines).  The arithmetic (DC interpolation, normalisation,
// dB, baseline) runs on the device: rpf_write_data().
void Acquisition::write_data(std::ostream &out) const
{
    const int N = params.N;
    if (params.matrixMode) {
        // one scan row: N float32 values appended to the .bin file (src/acquisition.cxx:385-388,400-405);
        // normalisation, dB and baseline run on the device, narrowed to float there (fpwrdb = pwrdb, :404)
        std::vector<float> row(N);
        int rc = rpf_write_row_f32(data.context(), actual_samplerate, params.linear ? 1 : 0,
                                   params.baseline ? aux.baseline_values.data() : nullptr, row.data());
        if (rc != RPF_OK)
            throw RPFexception(std::string("write_data: ") + rpf_strerror(data.context()), ReturnValue::HardwareError);
        data.pwr[N / 2] = (data.pwr[N / 2 - 1] + data.pwr[(N / 2 + 1) % N]) / 2;   // :377, as below
        std::ofstream binfile(params.bin_file, std::ios::out | std::ios::app | std::ios::binary);
        if (!binfile.good()) throw RPFexception("Could not open " + params.bin_file + ".", ReturnValue::InvalidInput);
        binfile.write(reinterpret_cast<const char *>(row.data()), (std::streamsize)sizeof(float) * N);
        return;
    }
    std::vector<double> f(N), v(N);
    int rc = rpf_write_data(data.context(), (double)tuned_freq, actual_samplerate, params.linear ? 1 : 0,
                            params.baseline ? aux.baseline_values.data() : nullptr, f.data(), v.data());
    if (rc != RPF_OK)
        throw RPFexception(std::string("write_data: ") + rpf_strerror(data.context()), ReturnValue::HardwareError);
    // the reference mutates pwr[N/2] while printing (src/acquisition.cxx:377); keep that visible
    data.pwr[N / 2] = (data.pwr[N / 2 - 1] + data.pwr[(N / 2 + 1) % N]) / 2;   // (N = 2: the reference reads past the end)

    out << "# rtl-power-fftw output" << std::endl;
    out << "#" << std::endl;
    out << "# frequency [Hz] power spectral density [dB/Hz]" << std::endl;
    const int extraDigitsFreq = 2;
    // src/acquisition.cxx:381-382: actual_samplerate / params.N is an INTEGER division there (bin width truncated to
    // whole Hz); with N > samplerate it is 0 and the expression overflows to a negative precision, which iostreams
    // treat as the default 6
    const int binHz = actual_samplerate / N;
    const int significantPlacesFreq = binHz > 0
        ? (int)std::ceil(std::floor(std::log10((double)tuned_freq)) - std::log10((double)binHz) + 1 + extraDigitsFreq)
        : 6;
    const int significantPlacesPwr = 6;
    for (int i = 0; i < N; i++)
        out << std::setprecision(significantPlacesFreq) << f[i] << " " << std::setprecision(significantPlacesPwr)
            << v[i] << std::endl;
    out << std::endl;
    out.flush();
}
