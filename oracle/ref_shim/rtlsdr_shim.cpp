// rtlsdr_shim.cpp -- TEST INFRASTRUCTURE ONLY: a fake RTL-SDR dongle behind oracle/ref_shim/rtl-sdr.h.
// One device; rtlsdr_read_sync hands out consecutive bytes of the stream given to rpf_shim_rtlsdr_set_stream
// (wrapping around at its end); tuning to a new centre frequency rewinds the stream, so every hop sees the same
// bytes.  Stands in for the USB hardware the reference's src/device.cxx drives.
#include <string.h>

#include "rtl-sdr.h"

struct rtlsdr_dev {
    uint32_t freq, rate;
    uint64_t pos;
};
static rtlsdr_dev g_dev = {0, 0, 0};
static const uint8_t *g_bytes = nullptr;
static uint64_t g_len = 0;

extern "C" {
void rpf_shim_rtlsdr_set_stream(const uint8_t *bytes, uint64_t len) { g_bytes = bytes; g_len = len; }
uint32_t rtlsdr_get_device_count(void) { return 1; }
int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t) { *dev = &g_dev; g_dev.pos = 0; return 0; }
int rtlsdr_close(rtlsdr_dev_t *) { return 0; }
int rtlsdr_get_tuner_gains(rtlsdr_dev_t *, int *gains) { if (gains) gains[0] = 372; return 1; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *, int) { return 0; }
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *, int) { return 0; }
int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq) { dev->freq = freq; dev->pos = 0; return 0; }
uint32_t rtlsdr_get_center_freq(rtlsdr_dev_t *dev) { return dev->freq; }
int rtlsdr_set_freq_correction(rtlsdr_dev_t *, int) { return 0; }
int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate) { dev->rate = rate; return 0; }
uint32_t rtlsdr_get_sample_rate(rtlsdr_dev_t *dev) { return dev->rate; }
int rtlsdr_reset_buffer(rtlsdr_dev_t *) { return 0; }
int rtlsdr_read_sync(rtlsdr_dev_t *dev, void *buf, int len, int *n_read)
{
    uint8_t *out = (uint8_t *)buf;
    int done = 0;
    while (done < len && g_len) {
        const uint64_t off = dev->pos % g_len;
        uint64_t n = g_len - off;
        if (n > (uint64_t)(len - done)) n = (uint64_t)(len - done);
        memcpy(out + done, g_bytes + off, n);
        done += (int)n;
        dev->pos += n;
    }
    *n_read = done;
    return 0;
}
}
