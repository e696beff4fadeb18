// rpf_power_synth -- the reference's main() (src/rtl_power_fftw.cxx:107-205) fed by the synthetic source
// instead of an RTL-SDR dongle:
//   Plan -> Datastore -> for every centre frequency: Acquisition::run -> print_summary -> write_data -> histogram.
// usage: rpf_power_synth N repeats [window:0|1] [seed] [mode] [buf_length|0] [buffers] [linear:0|1] [device]
//                        [-w window_file|-] [-B baseline_file|-] [--dump-aux] [--hops H]
//                        [--rate Hz] [--integration-time SECONDS] [--freq-range START:STOP] [--overlap PERCENT]
//                        [--plan-only] [--matrix FILE.bin] [--devices 0,1,..|A-B] [--pipelined] [--hop-parallel]
//   --devices      ONE Datastore spanning these GPUs (rpf_config.devices[]): every acquisition's spectra are dealt
//                  over them and summed by one NCCL allreduce.
//   --hop-parallel with --devices: one single-GPU Datastore per GPU instead, and the centre frequencies of the scan
//                  are dealt over them (hop h on GPU h mod n), output in frequency order.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <thread>

#include "acquisition.h"

static std::vector<int> parse_devices(const std::string &a)
{
    std::vector<int> out;
    const size_t dash = a.find('-');
    if (dash != std::string::npos && a.find(',') == std::string::npos) {
        const int lo = std::atoi(a.substr(0, dash).c_str()), hi = std::atoi(a.substr(dash + 1).c_str());
        for (int d = lo; d <= hi; d++) out.push_back(d);
        return out;
    }
    std::stringstream ss(a);
    std::string tok;
    while (std::getline(ss, tok, ',')) if (!tok.empty()) out.push_back(std::atoi(tok.c_str()));
    return out;
}

int main(int argc, char **argv)
{
    try {
        if (argc < 3) {
            std::cerr << "usage: " << argv[0]
                      << " N repeats [window] [seed] [mode] [buf_length] [buffers] [linear] [device] [options]" << std::endl;
            return (int)ReturnValue::InvalidArgument;
        }
        Params params;
        params.N = std::atoi(argv[1]);
        if (params.N % 2 != 0) params.N++;                         // src/params.cxx:152-155
        params.repeats = std::atoll(argv[2]);
        // positional arguments end at the first one that starts with '-' (options may follow directly)
        int npos = 3;
        while (npos < argc && npos < 10 && argv[npos][0] != '-') npos++;
        auto pos = [&](int i) { return i < npos ? argv[i] : nullptr; };
        params.window = pos(3) && std::atoi(pos(3)) != 0;
        const uint64_t seed = pos(4) ? std::strtoull(pos(4), nullptr, 0) : 0x5EED;
        const int mode = pos(5) ? std::atoi(pos(5)) : 1;
        if (pos(6) && std::atoi(pos(6)) > 0) { params.buf_length = std::atoi(pos(6)); params.buf_length_isSet = true; }
        if (pos(7)) params.buffers = std::atoi(pos(7));
        params.linear = pos(8) && std::atoi(pos(8)) != 0;
        if (pos(9)) params.cuda_device = std::atoi(pos(9));
        bool dump_aux = false, plan_only = false, hop_parallel = false;
        int hops = 1;                                              // acquisitions on ONE Datastore, a fresh worker each
        for (int i = npos; i < argc; i++) {
            std::string a = argv[i];
            auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
            if (a == "-w") { params.window = true; params.window_file = next(); }
            else if (a == "-B") { params.baseline = true; params.baseline_file = next(); }
            else if (a == "--dump-aux") dump_aux = true;
            else if (a == "--hops") hops = std::atoi(next());
            else if (a == "--rate") params.sample_rate = std::atoi(next());
            else if (a == "--integration-time") { params.integration_time = std::atof(next()); params.integration_time_isSet = true; }
            else if (a == "--freq-range") {
                std::string r = next();
                const size_t c = r.find(':');
                if (c == std::string::npos) throw RPFexception("--freq-range START:STOP", ReturnValue::InvalidArgument);
                params.startfreq = std::atoll(r.substr(0, c).c_str());
                params.stopfreq = std::atoll(r.substr(c + 1).c_str());
                params.freq_hopping_isSet = true;
            }
            else if (a == "--overlap") params.min_overlap = std::atof(next());
            else if (a == "--plan-only") plan_only = true;
            else if (a == "--matrix") { params.matrixMode = true; params.bin_file = next(); }
            else if (a == "--devices") params.cuda_devices = parse_devices(next());
            else if (a == "--pipelined") params.pipelined_push = true;
            else if (a == "--hop-parallel") hop_parallel = true;
            else throw RPFexception("unknown option " + a, ReturnValue::InvalidArgument);
        }
        // stdout is the data channel (gnuplot text / nothing in matrix mode): keep NCCL's own log lines off it
        if (params.cuda_devices.size() > 1) setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        if (params.buf_length % base_buf != 0)                     // src/params.cxx:166-170
            params.buf_length = (int)(std::floor((double)params.buf_length / base_buf + 0.5) * base_buf);

        AuxData aux;
        if (!params.window_file.empty() || !params.baseline_file.empty()) {
            aux = AuxData(params);                                  // text files / stdin, as the reference's AuxData
            if (dump_aux) {                                         // parser check, no GPU involved
                std::cout.precision(17);
                std::cout << "window " << aux.window_values.size() << std::endl;
                for (float v : aux.window_values) std::cout << v << std::endl;
                std::cout << "baseline " << aux.baseline_values.size() << std::endl;
                for (double v : aux.baseline_values) std::cout << v << std::endl;
                return (int)ReturnValue::Success;
            }
        } else if (params.window) {                                // Hann, as BASELINE config 3
            aux.window_values.resize(params.N);
            for (int n = 0; n < params.N; n++)
                aux.window_values[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * n / params.N));
        }
        Plan plan(params, params.sample_rate);
        if (plan_only) {                                           // planner check, no GPU involved
            std::cout << "repeats " << params.repeats << std::endl;
            std::cout << "buf_length " << params.buf_length << std::endl;
            std::cout << "freqs " << plan.freqs_to_tune.size();
            for (int64_t f : plan.freqs_to_tune) std::cout << " " << f;
            std::cout << std::endl;
            plan.print(std::cerr);
            return (int)ReturnValue::Success;
        }
        if (params.matrixMode) {                                   // src/rtl_power_fftw.cxx:124-129: truncate, then append rows
            FILE *f = std::fopen(params.bin_file.c_str(), "wb");
            if (!f) throw RPFexception("Could not open " + params.bin_file + ".", ReturnValue::InvalidInput);
            std::fclose(f);
        }
        params.finalfreq = plan.freqs_to_tune.back();
        std::vector<int64_t> freqs;
        for (int pass = 0; pass < hops; pass++)                    // --hops H: the frequency list H times over
            freqs.insert(freqs.end(), plan.freqs_to_tune.begin(), plan.freqs_to_tune.end());
        const bool per_freq = params.freq_hopping_isSet;

        if (hop_parallel && params.cuda_devices.size() > 1) {
            // One single-GPU Datastore per GPU; hop h runs on GPU h mod n in its own thread; the text blocks and the
            // stderr summaries are collected and emitted in frequency order (matrix rows likewise).
            const std::vector<int> devs = params.cuda_devices;
            const size_t nd = devs.size();
            std::vector<std::string> out(freqs.size()), err(freqs.size());
            std::vector<std::string> failure(nd);
            std::vector<int> failcode(nd, 0);
            std::vector<std::thread> th;
            for (size_t g = 0; g < nd; g++) {
                th.emplace_back([&, g]() {
                    try {
                        Params p = params;
                        p.cuda_devices.clear();
                        p.cuda_device = devs[g];
                        p.matrixMode = false;                      // rows are written by the main thread, in order
                        Datastore data(p, aux.window_values);
                        SyntheticSource source(seed, mode, per_freq);
                        for (size_t h = g; h < freqs.size(); h += nd) {
                            Acquisition acq(p, aux, source, data, p.sample_rate, freqs[h]);
                            acq.run();
                            std::ostringstream o, e;
                            acq.print_summary(e);
                            if (params.matrixMode) {
                                std::vector<float> row(p.N);
                                if (rpf_write_row_f32(data.context(), p.sample_rate, p.linear ? 1 : 0,
                                                      p.baseline ? aux.baseline_values.data() : nullptr, row.data()) != RPF_OK)
                                    throw RPFexception(std::string("write_data: ") + rpf_strerror(data.context()), ReturnValue::HardwareError);
                                o.write(reinterpret_cast<const char *>(row.data()), (std::streamsize)sizeof(float) * p.N);
                            } else {
                                acq.write_data(o);
                            }
                            out[h] = o.str();
                            err[h] = e.str();
                        }
                    } catch (const RPFexception &ex) {
                        failure[g] = ex.what();
                        failcode[g] = (int)ex.returnValue();
                    }
                });
            }
            for (auto &t : th) t.join();
            for (size_t g = 0; g < nd; g++)
                if (failcode[g]) throw RPFexception(failure[g], (ReturnValue)failcode[g]);
            FILE *bin = params.matrixMode ? std::fopen(params.bin_file.c_str(), "ab") : nullptr;
            for (size_t h = 0; h < freqs.size(); h++) {
                std::cerr << err[h];
                if (bin) std::fwrite(out[h].data(), 1, out[h].size(), bin);
                else std::cout << out[h];
            }
            if (bin) std::fclose(bin);
            return (int)ReturnValue::Success;
        }

        SyntheticSource source(seed, mode, per_freq);
        Datastore data(params, aux.window_values);
        for (int64_t f : freqs) {                                  // the hop loop of src/rtl_power_fftw.cxx:132-205
            Acquisition acquisition(params, aux, source, data, params.sample_rate, f);
            acquisition.run();
            acquisition.print_summary();
            acquisition.write_data(std::cout);
            data.printQueueHistogram();
        }
        return (int)ReturnValue::Success;
    } catch (const RPFexception &e) {
        std::cerr << e.what() << std::endl;
        return (int)e.returnValue();
    }
}
