/* wn_mock.h -- controls of the CPU stand-in for the C ABI (tests/mock/wn_mock_abi.cpp; TEST INFRASTRUCTURE ONLY). */
#ifndef WN_MOCK_H
#define WN_MOCK_H
#ifdef __cplusplus
extern "C" {
#endif
/* every batch call on `device` sleeps us_per_frame * n_frames afterwards (a GPU behind a shared PCIe uplink) */
void wn_mock_set_delay_us(int device, long us_per_frame);
/* batch calls made on `device` so far */
long wn_mock_batch_calls(int device);
/* batch calls on `device` fail with WN_ERR_CUDA after `after_calls` more successful ones (device < 0: never) */
void wn_mock_fail_batches(int device, long after_calls);
/* pinned allocations on `device` fail with WN_ERR_OOM after `after_allocs` more successful ones (device < 0: never) */
void wn_mock_fail_pinned_alloc(int device, long after_allocs);
#ifdef __cplusplus
}
#endif
#endif
