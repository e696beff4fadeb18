// rpf_power_synth -- the reference's main() hop body (src/rtl_power_fftw.cxx:107-170)
// for ONE hop, fed by the synthetic source instead of an RTL-SDR dongle:
//   Plan -> Datastore -> Acquisition::run -> print_summary -> write_data -> histogram.
// usage: rpf_power_synth N repeats [window:0|1] [seed] [mode] [buf_length|0] [buffers] [linear:0|1] [device]
//                        [-w window_file|-] [-B baseline_file|-] [--dump-aux] [--hops H]
#include <cmath>
#include <cstdlib>
#include <iostream>

#include "acquisition.h"

int main(int argc, char **argv)
{
    try {
        if (argc < 3) {
            std::cerr << "usage: " << argv[0]
                      << " N repeats [window] [seed] [mode] [buf_length] [buffers] [linear] [device]" << std::endl;
            return (int)ReturnValue::InvalidArgument;
        }
        Params params;
        params.N = std::atoi(argv[1]);
        if (params.N % 2 != 0) params.N++;                         // src/params.cxx:152-155
        params.repeats = std::atoll(argv[2]);
        params.window = argc > 3 && std::atoi(argv[3]) != 0;
        const uint64_t seed = argc > 4 ? std::strtoull(argv[4], nullptr, 0) : 0x5EED;
        const int mode = argc > 5 ? std::atoi(argv[5]) : 1;
        if (argc > 6 && std::atoi(argv[6]) > 0) { params.buf_length = std::atoi(argv[6]); params.buf_length_isSet = true; }
        if (argc > 7) params.buffers = std::atoi(argv[7]);
        params.linear = argc > 8 && std::atoi(argv[8]) != 0;
        if (argc > 9) params.cuda_device = std::atoi(argv[9]);
        bool dump_aux = false;
        int hops = 1;                                              // acquisitions on ONE Datastore, a fresh worker each
        for (int i = 10; i < argc; i++) {                          // optional: -w FILE  -B FILE  --dump-aux
            std::string a = argv[i];
            if (a == "-w" && i + 1 < argc) { params.window = true; params.window_file = argv[++i]; }
            else if (a == "-B" && i + 1 < argc) { params.baseline = true; params.baseline_file = argv[++i]; }
            else if (a == "--dump-aux") dump_aux = true;
            else if (a == "--hops" && i + 1 < argc) hops = std::atoi(argv[++i]);
        }
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
        SyntheticSource source(seed, mode);
        Plan plan(params, params.sample_rate);
        Datastore data(params, aux.window_values);
        for (int hop = 0; hop < hops; hop++) {                     // the hop loop of src/rtl_power_fftw.cxx:132-205
            source.rewind();                                       // same bytes every hop: outputs must repeat exactly
            Acquisition acquisition(params, aux, source, data, params.sample_rate, params.cfreq);
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
