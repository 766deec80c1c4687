// Training step (forward with saved activations, BPTT, AdamW) -- placeholder until lstm_train.cu lands.
#include "lstm_handle.cuh"
namespace bci {
int lstm_forward_train(bci_lstm_s*, const float*, int, int, float, uint64_t, float*, float*, float*, void*, size_t, cudaStream_t) {
  set_error("bci_lstm_forward(train=1): not implemented in this build");
  return BCI_EINVAL;
}
size_t lstm_workspace_train(const bci_lstm_config&, int, int) { return 0; }
int lstm_backward_impl(bci_lstm_s*, const float*, const float*, int, int, float*, const bci_lstm_grads*, void*, size_t, cudaStream_t) {
  set_error("bci_lstm_backward: not implemented in this build");
  return BCI_EINVAL;
}
}  // namespace bci
extern "C" int bci_adamw_step(float*, const float*, float*, float*, int64_t, float, float, float, float, float, int32_t, float,
                              float, float*, void*) {
  bci::set_error("bci_adamw_step: not implemented in this build");
  return BCI_EINVAL;
}
