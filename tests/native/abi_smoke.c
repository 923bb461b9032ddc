/* A plain-C host of the C ABI (no C++, no CUDA headers, no Python): links libegm_b200.so, checks the
 * version, runs the pure-host size queries and the argument validation. Needs no GPU.
 *   gcc -std=c99 -I include tests/native/abi_smoke.c -L ego-moment-cle-vit_b200/lib -legm_b200 -o abi_smoke */
#include <stdio.h>
#include <string.h>

#include "egm_b200.h"

int main(void) {
  const int B = 256, N = 197, D = 768, K = 5;
  if (egm_version() < 100) return 1;
  if (egm_gpf_ldr(N) != 200) return 2;
  /* 13 D x D operand-plane matrices of state for 5 Newton-Schulz iterations */
  size_t ns = egm_ns_state_bytes(B, D, K, EGM_PREC_BF16X3);
  if (ns < (size_t)13 * B * D * D * 4) return 3;
  size_t mhd = egm_mhd_state_bytes(B, N, D, K, EGM_PREC_BF16X3);
  if (mhd <= ns) return 4;
  /* a null pointer is an argument error with a message, never a crash */
  int rc = egm_triu_pack(NULL, B, D, NULL, NULL);
  if (rc != EGM_ERR_ARG || strstr(egm_last_error(), "egm_triu_pack") == NULL) return 5;
  rc = egm_mhd_fwd(NULL, NULL, B, N, D, K, 1e-5f, EGM_MHD_SYMMETRIC_GRAPH, NULL, NULL, NULL, NULL, NULL, NULL,
                   EGM_PREC_BF16X3, NULL, 0, NULL);
  if (rc != EGM_ERR_ARG) return 6;
  printf("abi ok: version %d, NS state %zu MB, fused-head state %zu MB\n", egm_version(), ns >> 20, mhd >> 20);
  return 0;
}
