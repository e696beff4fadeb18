// ref_main.cpp -- TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/ref_rtl_power: the reference's OWN
// src/acquisition.cxx (AuxData, Plan, Acquisition::run / print_summary / write_data), src/device.cxx,
// src/datastore.cxx and src/interrupts.cxx, compiled unmodified from /root/reference, behind a fake dongle
// (rtlsdr_shim.cpp), the stand-in FFT (shim_fft.cpp) and this small main().  The reference's real main()
// (src/rtl_power_fftw.cxx) cannot be built -- its Params constructor needs TCLAP -- so this file provides the
// three things that live there: the constructor of Params (a no-op: the member initialisers of src/params.h are
// the defaults), the matrix-mode globals of src/metadata.h, and the hop loop of src/rtl_power_fftw.cxx:132-170.
// usage: ref_rtl_power STREAM_FILE N repeats [options as rpf_power_synth: -w -B --rate --integration-time
//        --freq-range --overlap --buf-length --buffers --linear --matrix --hops --plan-only]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "acquisition.h"
#include "exceptions.h"
#include "interrupts.h"
#include "metadata.h"

Params::Params(int, char **) {}

int metaCols = 0, metaRows = 1;
int startFreq = 0, endFreq = 0, stepFreq = 0;
time_t scanEnd, scanBeg;
float avgScanDur = 0, sumScanDur = 0;
std::string firstAcqTimestamp, lastAcqTimestamp;
int cntTimeStamps = 0;

int main(int argc, char **argv)
{
    if (argc < 4) { std::cerr << "usage: ref_rtl_power STREAM N repeats [options]" << std::endl; return 3; }
    try {
        std::ifstream sf(argv[1], std::ios::binary);
        std::vector<uint8_t> stream((std::istreambuf_iterator<char>(sf)), std::istreambuf_iterator<char>());
        rpf_shim_rtlsdr_set_stream(stream.data(), stream.size());
        Params params(0, nullptr);
        params.N = std::atoi(argv[2]);
        if (params.N % 2 != 0) params.N++;
        params.repeats = std::atoll(argv[3]);
        params.talkless = false;
        bool plan_only = false;
        int hops = 1;
        for (int i = 4; i < argc; i++) {
            std::string a = argv[i];
            auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
            if (a == "-w") { params.window = true; params.window_file = next(); }
            else if (a == "-B") { params.baseline = true; params.baseline_file = next(); }
            else if (a == "--rate") params.sample_rate = std::atoi(next());
            else if (a == "--integration-time") { params.integration_time = std::atof(next()); params.integration_time_isSet = true; }
            else if (a == "--freq-range") {
                std::string r = next();
                const size_t c = r.find(':');
                params.startfreq = std::atoll(r.substr(0, c).c_str());
                params.stopfreq = std::atoll(r.substr(c + 1).c_str());
                params.freq_hopping_isSet = true;
            }
            else if (a == "--overlap") params.min_overlap = std::atof(next());
            else if (a == "--buf-length") { params.buf_length = std::atoi(next()); params.buf_length_isSet = true; }
            else if (a == "--buffers") params.buffers = std::atoi(next());
            else if (a == "--linear") params.linear = true;
            else if (a == "--matrix") { params.matrixMode = true; params.bin_file = next(); }
            else if (a == "--hops") hops = std::atoi(next());
            else if (a == "--plan-only") plan_only = true;
            else { std::cerr << "unknown option " << a << std::endl; return 3; }
        }
        AuxData auxData(params);
        Rtlsdr rtldev(params.dev_index);
        rtldev.set_sample_rate(params.sample_rate);
        int actual_samplerate = rtldev.sample_rate();
        Plan plan(params, actual_samplerate);
        if (plan_only) {
            std::cout << "repeats " << params.repeats << std::endl;
            std::cout << "buf_length " << params.buf_length << std::endl;
            std::cout << "freqs " << plan.freqs_to_tune.size();
            for (auto f : plan.freqs_to_tune) std::cout << " " << f;
            std::cout << std::endl;
            plan.print();
            return 0;
        }
        Datastore data(params, auxData.window_values);
        if (params.matrixMode) { std::ofstream t(params.bin_file, std::ios::out | std::ios::trunc | std::ios::binary); }
        params.finalfreq = plan.freqs_to_tune.back();
        for (int pass = 0; pass < hops; pass++)
            for (auto f : plan.freqs_to_tune) {
                Acquisition acquisition(params, auxData, rtldev, data, actual_samplerate, f);
                acquisition.run();
                acquisition.print_summary();
                acquisition.write_data();
                data.printQueueHistogram();
            }
        return 0;
    } catch (RPFexception &e) {
        std::cerr << e.what() << std::endl;
        return (int)e.returnValue();
    }
}
