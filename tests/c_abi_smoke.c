/* Plain-C consumer of include/pwv.h: proves the boundary is a C ABI (no C++/torch types) and
 * exercises the host-side validation paths that need no GPU. Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "../include/pwv.h"

int main(void) {
  pwv_hparams hp;
  pwv_model* m = NULL;
  int i, j, n;
  const char* name = NULL;
  int64_t shape[4];
  int ndim = 0;
  float w[80 * 80];

  memset(&hp, 0, sizeof hp);
  hp.n_iaf = 2; hp.filter_width = 2;
  hp.residual_channels = 64; hp.dilation_channels = 64; hp.skip_channels = 128;
  hp.condition_channels = 80; hp.n_mels = 80; hp.hop_length = 80;
  hp.use_biases = 1; hp.use_skip_connection = 0; hp.precision = PWV_PREC_F16X3;
  for (i = 0; i < 2; ++i) {
    hp.n_layers[i] = 3;
    for (j = 0; j < 3; ++j) hp.dilations[i][j] = 1 << j;
  }
  if (pwv_version() != PWV_VERSION) return 1;
  if (pwv_model_create(&hp, &m) != PWV_OK || !m) { printf("create: %s\n", pwv_last_error()); return 2; }
  n = pwv_model_num_variables(m);
  if (n != 1 + 2 * 2 * (1 + 3 * 10 + 4)) { printf("unexpected variable count %d\n", n); return 3; }
  if (pwv_model_variable(m, 0, &name, shape, &ndim) != PWV_OK || strcmp(name, "iaf_vocoder/cond/dense") != 0 || ndim != 3 ||
      shape[0] != 1 || shape[1] != 80 || shape[2] != 80) return 4;
  memset(w, 0, sizeof w);
  if (pwv_model_load_weight(m, name, w, shape, ndim) != PWV_OK) return 5;
  if (pwv_model_load_weight(m, "no/such/variable", w, shape, ndim) != PWV_ENAME) return 6;
  if (pwv_model_finalize(m) != PWV_ESTATE) return 7;              /* not every variable was loaded */
  if (strstr(pwv_last_error(), "was not loaded") == NULL) return 8;
  {
    size_t bytes = 0;
    if (pwv_workspace_bytes(m, 8, 16000, &bytes) != PWV_OK || bytes < (size_t)2 * 2 * 8 * 16000 * 64 * 4) return 9;
    if (pwv_workspace_bytes(m, 8, 16001, &bytes) != PWV_EINVAL) return 10;
  }
  hp.filter_width = 3;
  {
    pwv_model* bad = NULL;
    if (pwv_model_create(&hp, &bad) != PWV_EINVAL || bad != NULL) return 11;
  }
  if (pwv_model_destroy(m) != PWV_OK) return 12;
  printf("c_abi_smoke ok (%d variables)\n", n);
  return 0;
}
