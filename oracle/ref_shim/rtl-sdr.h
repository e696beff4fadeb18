/*
 * rtl-sdr.h -- SHIM, test infrastructure only (see oracle/README.md).
 *
 * The reference's src/device.{h,cxx} and src/acquisition.{h,cxx} include <rtl-sdr.h> (librtlsdr, not installed
 * here).  This header declares exactly the librtlsdr entry points src/device.cxx calls (src/device.cxx:31,44,53,
 * 57,64,73,83,94,95,101,102,112,125,133) so that the UNMODIFIED reference sources compile from /root/reference
 * into oracle/_ref/.  rtlsdr_shim.cpp defines them as a fake dongle that replays a byte stream from memory.
 */
#ifndef RPF_ORACLE_RTLSDR_SHIM_H
#define RPF_ORACLE_RTLSDR_SHIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct rtlsdr_dev rtlsdr_dev_t;
uint32_t rtlsdr_get_device_count(void);
int rtlsdr_open(rtlsdr_dev_t **dev, uint32_t index);
int rtlsdr_close(rtlsdr_dev_t *dev);
int rtlsdr_get_tuner_gains(rtlsdr_dev_t *dev, int *gains);
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *dev, int manual);
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *dev, int gain);
int rtlsdr_set_center_freq(rtlsdr_dev_t *dev, uint32_t freq);
uint32_t rtlsdr_get_center_freq(rtlsdr_dev_t *dev);
int rtlsdr_set_freq_correction(rtlsdr_dev_t *dev, int ppm);
int rtlsdr_set_sample_rate(rtlsdr_dev_t *dev, uint32_t rate);
uint32_t rtlsdr_get_sample_rate(rtlsdr_dev_t *dev);
int rtlsdr_reset_buffer(rtlsdr_dev_t *dev);
int rtlsdr_read_sync(rtlsdr_dev_t *dev, void *buf, int len, int *n_read);

/* shim control (not librtlsdr): the byte stream the fake dongle replays; every retune restarts it */
void rpf_shim_rtlsdr_set_stream(const uint8_t *bytes, uint64_t len);
#ifdef __cplusplus
}
#endif
#endif
