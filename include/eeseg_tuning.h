/* Tuning-only entry points of libeeseg_b200_tuning.so (built with -DEESEG_TUNING by
 * `python -m ee_semantic_segmentation_b200.build --tuning`). NOT part of the product C ABI (include/eeseg.h): the
 * shipped library contains neither these symbols nor the in-kernel instrumentation behind them. They keep
 * process-global, unsynchronised state and are meant for single-threaded measurement scripts (tools/conv_debug.py,
 * tools/conv_step_times.py). */
#ifndef EESEG_TUNING_H_
#define EESEG_TUNING_H_
#ifdef __cplusplus
extern "C" {
#endif

/* Following conv launches record {first CTA start, last CTA end} in wall-clock ns (%globaltimer) at buffer[2*i],
 * buffer[2*i+1] (uint64, i < capacity; initialise starts to UINT64_MAX and ends to 0). NULL switches it off; returns
 * how many launches were recorded. */
int eeseg_conv_timing(void* device_buffer, int capacity);

/* Device buffer of [148][32] uint64 cycle counters (per-CTA wait times of the producer, MMA and epilogue roles)
 * filled by subsequent conv launches; NULL switches it off. */
int eeseg_conv_debug_stats(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* EESEG_TUNING_H_ */
