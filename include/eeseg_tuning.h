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

/* Timing probes of the conv kernel: skip_mask bit 0 = no activation tiles, 1 = no weight tiles, 2 = no residual tiles,
 * 3 = no output stores (the launches' outputs are then garbage), 4 = register -> global epilogue on every multi-tile
 * launch, 5 = operand ring capped at two stages, 7 = the un-pipelined epilogue column loop; max_ctas > 0 caps the grid.
 * (0, 0) restores normal launches. */
int eeseg_conv_probe(int skip_mask, int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* EESEG_TUNING_H_ */
