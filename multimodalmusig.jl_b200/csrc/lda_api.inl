// lda_api.inl -- host side of the LDA entry points (included by mmsig_api.cu)
#define LDA_TODO return fail(h, MMSIG_EINVAL, "LDA path not built yet")
extern "C" int32_t mmsig_lda_set_data(mmsig_handle *h, int64_t, int64_t, int32_t, int32_t, const int64_t *, const int32_t *, const int32_t *) { LDA_TODO; }
extern "C" int32_t mmsig_lda_set_state(mmsig_handle *h, double, double, const double *, const double *) { LDA_TODO; }
extern "C" int32_t mmsig_lda_iterate(mmsig_handle *h, double *) { LDA_TODO; }
extern "C" int32_t mmsig_lda_fit(mmsig_handle *h, int32_t, double, double *, int32_t *, int32_t *) { LDA_TODO; }
extern "C" int32_t mmsig_lda_elbo(mmsig_handle *h, double *, double *) { LDA_TODO; }
extern "C" int32_t mmsig_lda_get_state(mmsig_handle *h, double *, double *, double *, double *, double *, double *) { LDA_TODO; }
extern "C" int32_t mmsig_lda_get_phi(mmsig_handle *h, double *) { LDA_TODO; }
