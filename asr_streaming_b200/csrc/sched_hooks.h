// Engine entry points used by the native session scheduler (sched.cu); not part of the C ABI.
#pragma once
#include <stdint.h>

struct AsrEngine;

namespace asr {

struct StepView {          // pointers into the pinned result area of a collected ticket: valid until two more steps were submitted
  int n = 0;
  const int32_t *argmax_ids = nullptr, *new_tokens = nullptr, *n_new = nullptr, *blank_frames = nullptr, *has_token = nullptr, *has_text = nullptr, *flags = nullptr;
  const int16_t* beam_tokens = nullptr;     // [n][ASR_BEAM_MAX_LEN] or null (no beam)
  const int32_t* beam_len = nullptr;
  const float* beam_score = nullptr;
  const float* logprobs = nullptr;          // null unless requested at submit
};

// Batch assembly + submit in one call: chunk i = base[rows[i] * row_stride + offsets[i] ..) — host gather (multi-threaded memcpy into the
// pinned staging buffer of the step) or device gather (`base` pinned + mapped: a kernel reads the chunks over PCIe).
int engine_submit_gather(AsrEngine* e, int n, const int32_t* slots, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets,
                         bool device_gather, bool want_logprobs, int* ticket);
// Pre-staging (see engine.cu): gather + H2D of candidate chunks ahead of the decision which of them run; then the step over a subset.
int engine_prestage(AsrEngine* e, int n_rows, const int16_t* base, int64_t row_stride, const int32_t* rows, const int64_t* offsets, bool device_gather);
int engine_submit_prestaged(AsrEngine* e, int n, const int32_t* slots, const int32_t* staged_index, bool want_logprobs, int* ticket);
int engine_collect_view(AsrEngine* e, int ticket, StepView* v);
int engine_reset_async(AsrEngine* e, int n, const int32_t* slots);      // asr_session_reset_many without the argument checks' error text
int engine_open_slot(AsrEngine* e, int32_t* slot);
int engine_close_slot(AsrEngine* e, int32_t slot);
int engine_wait_inputs(AsrEngine* e);

}  // namespace asr
