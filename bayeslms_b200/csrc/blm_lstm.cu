// placeholder until the persistent recurrence kernel lands (next commit)
#include "blm_host.h"
namespace blm { int lstm_init() { return BLM_OK; } }
extern "C" int64_t blm_lstm_workspace_bytes(int64_t, int64_t) { return 64; }
extern "C" int blm_lstm_layer(const float*, const blm_bf16*, const blm_bf16*, const float*, const float*,
                              const int32_t*, int64_t, int64_t, int64_t, float*, blm_bf16*, blm_bf16*, float*,
                              float*, void*, blm_stream) {
  blm::set_error("blm_lstm_layer: not built yet");
  return BLM_ERR_ARG;
}
