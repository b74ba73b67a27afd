// Bias denoiser (hifigan/denoiser.py): STFT -> magnitude minus strength*bias -> ISTFT.  Filled in below.
#include "ctx.cuh"

using namespace ev;

extern "C" size_t ev_denoise_workspace_bytes(const ev_ctx* ctx, int B, int L) {
  (void)ctx; (void)B; (void)L;
  return 0;
}
extern "C" int ev_denoiser_init(ev_ctx* ctx, float* bias_spec_out, void* workspace, size_t workspace_bytes, void* stream) {
  (void)bias_spec_out; (void)workspace; (void)workspace_bytes; (void)stream;
  return fail(ctx, EV_ERR_INVALID, "ev_denoiser_init: not implemented yet");
}
extern "C" int ev_denoise(ev_ctx* ctx, const float* audio, int B, int L, float strength, float* out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  (void)audio; (void)B; (void)L; (void)strength; (void)out; (void)workspace; (void)workspace_bytes; (void)stream;
  return fail(ctx, EV_ERR_INVALID, "ev_denoise: not implemented yet");
}
