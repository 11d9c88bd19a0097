/* A plain C99 client of include/vbsky_b200.h: no Python, no torch, host buffers only.
 * Tree of tree_data.py's doctest "((human:1,chimp:2):1,gorilla:2.5)": tips 0,1,2; node 3 = (0,1); root 4 = (3,2).
 * Prints the per-unit log-likelihood and d ll / d bl of two units; tests/test_cabi_client.py compares them with the
 * oracle.  Exit code 0 on success, 1 + message on any error. */
#include <stdio.h>
#include <stdlib.h>

#include "vbsky_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != VBSKY_OK) {                                                   \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, vbsky_last_error()); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  const int32_t parent_child[5][2] = {{-1, -1}, {-1, -1}, {-1, -1}, {1, 0}, {3, 2}};
  const int32_t child_parent[5] = {3, 3, 4, 4, -1};
  const int32_t postorder[5] = {0, 1, 3, 2, 4};
  const double sample_times[3] = {1.0, 0.0, 0.5};
  /* four site patterns x three tips, state masks A=1 C=2 G=4 T=8 N=15 */
  const uint8_t patterns[4][3] = {{1, 1, 1}, {1, 2, 1}, {4, 4, 8}, {15, 1, 2}};
  const double counts[4] = {1.0, 2.0, 1.0, 3.0};
  const int32_t unit_tree[2] = {0, 0};
  const double bl[2][4] = {{0.1, 0.2, 0.25, 0.05}, {0.3, 0.01, 0.02, 0.4}};
  double ll[2], dll[2][4];
  vbsky_subst_model model;
  vbsky_layout* lay = NULL;
  vbsky_tips* tips = NULL;
  vbsky_session* ses = NULL;
  int u, b;

  if (vbsky_device_count() < 1) {
    fprintf(stderr, "no CUDA device: %s\n", vbsky_last_error());
    return 1;
  }
  CHECK(vbsky_hky(2.7, &model));
  CHECK(vbsky_layout_create(1, 3, &parent_child[0][0], child_parent, postorder, sample_times, &lay));
  CHECK(vbsky_tips_create(lay, 4, &patterns[0][0], counts, &tips));
  CHECK(vbsky_session_create(lay, tips, 2, 0, &ses));
  CHECK(vbsky_prune_value_and_grad_host(ses, 2, unit_tree, &bl[0][0], &model, ll, &dll[0][0], NULL));
  for (u = 0; u < 2; ++u) {
    printf("%.17g", ll[u]);
    for (b = 0; b < 4; ++b) printf(" %.17g", dll[u][b]);
    printf("\n");
  }
  vbsky_session_destroy(ses);
  vbsky_tips_destroy(tips);
  vbsky_layout_destroy(lay);
  return 0;
}
