extern "C" size_t wealy_loss_workspace_bytes(int64_t, int64_t, int) { return 0; }
extern "C" int wealy_loss_forward(const wealy_loss_cfg*, const void*, int64_t, int64_t, int64_t, int, const int64_t*,
                                  const int64_t*, double*, void*, size_t, void*) {
  return fail(WEALY_ERR_UNSUPPORTED, "loss kernels not built yet");
}
extern "C" int wealy_loss_backward(const wealy_loss_cfg*, const void*, int64_t, int64_t, int64_t, int, const float*,
                                   void*, int64_t, void*, size_t, void*) {
  return fail(WEALY_ERR_UNSUPPORTED, "loss kernels not built yet");
}
