#include "lstm_handle.cuh"
namespace bci {
int lstm_forward_bf16(bci_lstm_s*, const float*, int, int, float*, float*, float*, void*, size_t, cudaStream_t) {
  set_error("bf16 path not built"); return BCI_EINVAL; }
size_t lstm_workspace_bf16(const bci_lstm_config&, int, int) { return 0; }
int lstm_pack_bf16(bci_lstm_s*, cudaStream_t) { return BCI_OK; }
size_t lstm_store_bytes_bf16(const bci_lstm_config&) { return 0; }
void lstm_carve_bf16(bci_lstm_s*, char*) {}
}
