// rpf_params.h -- the configuration, exit codes and exception type the hot
// path's callers use, with the reference's names and meanings so that code
// written against the reference's Datastore/Acquisition reads the same:
//   Params        <- src/params.h:29-67  (only the fields the worker, the
//                    producer and write_data read; no TCLAP constructor)
//   ReturnValue   <- src/exceptions.h:25-34 (same numeric exit codes)
//   RPFexception  <- src/exceptions.h:39-47
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

constexpr int base_buf = 16384;              // USB transfer quantum (src/params.h:26)
constexpr int default_buf_multiplier = 100;  // src/params.h:27

enum class ReturnValue : int {
    Success = 0, NoDeviceFound = 1, InvalidDeviceIndex = 2, InvalidArgument = 3,
    TCLAPerror = 4, InvalidInput = 5, AcquisitionError = 6, HardwareError = 7
};

class RPFexception : public std::runtime_error {
  public:
    RPFexception(const std::string &what, ReturnValue rv) : std::runtime_error(what), rv_(rv) {}
    ReturnValue returnValue() const { return rv_; }
  private:
    ReturnValue rv_;
};

struct Params {
    int N = 512;                                       // bins; forced even by the CLI
    int sample_rate = 2000000;
    int64_t cfreq = 1420405752;
    int buffers = 5;
    int buf_length = base_buf * default_buf_multiplier;
    bool buf_length_isSet = false;
    int64_t repeats = buf_length / (2 * 512);
    bool window = false;
    std::string window_file;                           // "-" = stdin
    bool baseline = false;
    std::string baseline_file;                         // "-" = stdin
    bool linear = false;
    bool talkless = true;
    int64_t startfreq = 0, stopfreq = 0;               // frequency hopping (src/params.h:36-37)
    bool freq_hopping_isSet = false;
    double min_overlap = 0;                            // percent (src/params.h:44)
    double integration_time = 0;                       // seconds (src/params.h:40-41)
    bool integration_time_isSet = false;
    bool matrixMode = false;                           // src/params.h:60-66: one float32 row per acquisition -> bin_file
    std::string bin_file;
    int64_t finalfreq = 0;
    int cuda_device = 0;                               // new: where the FFT worker runs
    std::vector<int> cuda_devices;                     // new: several GPUs behind ONE Datastore (rpf_config.devices[])
    bool pipelined_push = false;                       // new: RPF_FLAG_PIPELINED_PUSH, the worker holds one buffer back
};
