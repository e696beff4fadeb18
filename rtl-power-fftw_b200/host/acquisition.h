// acquisition.h -- the callers either side of the hot path, reference-shaped:
//   SampleSource  : what Acquisition needs of Rtlsdr (src/device.h:28-54):
//                   read(Buffer&), set_frequency(), frequency()
//   SyntheticSource: counter-based IQ bytes (rpf_synth_fill_host) instead of USB
//   AuxData       : window / baseline values (src/acquisition.h:33-41)
//   Plan          : repeats from the integration time, buffer auto-sizing, list of centre frequencies with
//                   overlap (src/acquisition.cxx:158-208)
//   Acquisition   : run() / print_summary() / write_data() (src/acquisition.h:81-98)
#pragma once
#include <cstdint>
#include <iosfwd>
#include <list>
#include <string>
#include <vector>

#include "datastore.h"

class SampleSource {
  public:
    virtual ~SampleSource() {}
    virtual bool read(Buffer &buffer) = 0;            // Rtlsdr::read (src/device.cxx:92-97)
    virtual void set_frequency(int64_t f) { freq = f; }
    virtual int64_t frequency() const { return freq; }
  protected:
    int64_t freq = 0;
};

class SyntheticSource : public SampleSource {
  public:
    // per_frequency: every centre frequency gets its own byte stream (seed mixed with the frequency), so the hops
    // of a scan differ; otherwise the same bytes are replayed at every frequency.
    SyntheticSource(uint64_t seed, int mode, bool per_frequency = false)
        : seed(seed), mode(mode), per_frequency(per_frequency) {}
    bool read(Buffer &buffer) override;
    void set_frequency(int64_t f) override { freq = f; offset = 0; }
    void rewind() { offset = 0; }
    uint64_t offset = 0;
  private:
    uint64_t seed;
    int mode;
    bool per_frequency;
};

// Window function and baseline read from text (src/acquisition.cxx:32-156): one value per line =
// the LAST number on the line, lines whose first non-blank character is '#' are comments, lines
// without a number are skipped; exactly N values are required (RPFexception, InvalidInput).
// Window values are float, baseline values double (src/acquisition.h:33-41).
class AuxData {
  public:
    AuxData() {}
    explicit AuxData(const Params &params);            // reads params.window_file / baseline_file
    AuxData(const Params &params, std::istream &in);   // as above with `in` standing for stdin ("-")
    std::vector<double> baseline_values;
    std::vector<float> window_values;
  private:
    void load(const Params &params, std::istream &in);
};

template <typename T> std::vector<T> read_inputfile(std::istream &stream);

class Plan {
  public:
    Plan(Params &params, int actual_samplerate);
    void print(std::ostream &err) const;              // src/acquisition.cxx:200-208
    std::list<int64_t> freqs_to_tune;                 // src/acquisition.h:53
    int actual_samplerate;
  protected:
    Params &params;
};

class Acquisition {
  public:
    Acquisition(const Params &params, AuxData &aux, SampleSource &source, Datastore &data,
                int actual_samplerate, int64_t freq);
    void run();
    void print_summary() const;
    void print_summary(std::ostream &err) const;
    // Text block on `out`, or -- params.matrixMode -- one float32 row appended to params.bin_file
    // (src/acquisition.cxx:385-388,400-405).
    void write_data(std::ostream &out) const;

    int64_t deviceReadouts = 0, successfulReadouts = 0;
  protected:
    const Params &params;
    AuxData &aux;
    SampleSource &source;
    Datastore &data;
    int actual_samplerate;
    int64_t freq;
    int64_t tuned_freq = 0;
};
