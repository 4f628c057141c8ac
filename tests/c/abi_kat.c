/* Plain-C consumer of include/mgn_b200.h: proves that the header is C (not C++), that the library links from C, and
 * runs the hand-derived known-answer tests of SURVEY.md section 8c through the host half of the ABI.  On a machine
 * without a GPU it also checks that the device entry points fail loudly (MGN_ERR_CUDA + message) instead of falling back.
 *   gcc -std=c11 -Wall -Werror -I include tests/c/abi_kat.c -o abi_kat -L meshgraphnets.jl_b200/csrc -lmgn_b200 -Wl,-rpath,...
 * Exit code 0 = all checks passed. */
#include <stdio.h>
#include <string.h>

#include "mgn_b200.h"

static int failures = 0;
#define CHECK(cond)                                                       \
  do {                                                                    \
    if (!(cond)) {                                                        \
      fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
      ++failures;                                                         \
    }                                                                     \
  } while (0)

static int eq_i32(const int32_t* a, const int32_t* b, int n) { return memcmp(a, b, sizeof(int32_t) * (size_t)n) == 0; }

int main(void) {
  char msg[256];
  CHECK(mgn_abi_version() == MGN_ABI_VERSION);

  /* one_hot([0,5,6], 7, 1): ones at rows 1, 6, 7 (1-based) */
  {
    const int32_t v[3] = {0, 5, 6};
    float out[21];
    CHECK(mgn_one_hot(v, 3, 7, 1, out) == MGN_OK);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 7; ++j) CHECK(out[i * 7 + j] == ((i == 0 && j == 0) || (i == 1 && j == 5) || (i == 2 && j == 6) ? 1.f : 0.f));
  }
  /* triangles_to_edges: faces (0,1,2),(1,2,3) -> U = 5, then the 0 -> 1 shift (src/graph.jl:30-34) */
  {
    const int32_t cells[6] = {0, 1, 2, 1, 2, 3};
    int32_t s[36], r[36];
    int64_t n = 0;
    int32_t shifted = 0;
    CHECK(mgn_triangles_to_edges(cells, 2, s, r, &n) == MGN_OK);
    CHECK(n == 10);
    const int32_t s0[10] = {1, 2, 3, 2, 3, 0, 1, 2, 0, 1}, r0[10] = {0, 1, 2, 0, 1, 1, 2, 3, 2, 3};
    CHECK(eq_i32(s, s0, 10) && eq_i32(r, r0, 10));
    CHECK(mgn_shift_one_based(s, r, n, &shifted) == MGN_OK && shifted == 1);
    const int32_t s1[10] = {2, 3, 4, 3, 4, 1, 2, 3, 1, 2}, r1[10] = {1, 2, 3, 1, 2, 2, 3, 4, 3, 4};
    CHECK(eq_i32(s, s1, 10) && eq_i32(r, r1, 10));
    CHECK(mgn_shift_one_based(s, r, n, &shifted) == MGN_OK && shifted == 0);   /* no zero left: no second shift */
  }
  /* parse_edges of the 1-D chain [i, i+1], N = 5 (src/dataset.jl:379-382, src/graph.jl:38) */
  {
    const int32_t e[8] = {1, 2, 2, 3, 3, 4, 4, 5};
    int32_t s[8], r[8];
    CHECK(mgn_parse_edges(e, 4, s, r) == MGN_OK);
    const int32_t s0[8] = {1, 2, 3, 4, 2, 3, 4, 5}, r0[8] = {2, 3, 4, 5, 1, 2, 3, 4};
    CHECK(eq_i32(s, s0, 8) && eq_i32(r, r0, 8));
  }
  /* edge features [rel ; |rel|] on a 3-4-5 triangle; out-of-range ids are an error with a message */
  {
    const float pos[6] = {0.f, 0.f, 3.f, 0.f, 3.f, 4.f};
    const int32_t s[2] = {3, 2}, r[2] = {1, 1};
    float out[6];
    CHECK(mgn_edge_features(pos, 3, 2, s, r, 2, 1, out) == MGN_OK);
    CHECK(out[0] == 3.f && out[1] == 4.f && out[2] == 5.f && out[3] == 3.f && out[4] == 0.f && out[5] == 3.f);
    const int32_t bad[2] = {4, 2};
    CHECK(mgn_edge_features(pos, 3, 2, bad, r, 2, 1, out) == MGN_ERR_INDEX);
    CHECK(mgn_last_error(msg, sizeof msg) == MGN_OK && strlen(msg) > 0);
  }
  /* argument errors never crash: status + message */
  CHECK(mgn_one_hot(NULL, 3, 7, 1, NULL) == MGN_ERR_INVALID);
  CHECK(mgn_model_create(NULL, NULL) == MGN_ERR_INVALID);
  {
    mgn_model_config cfg = {9, 3, 2, 64, 15, 2, 1e-5f, MGN_COMPUTE_BF16, 0, 0, 0};   /* bf16 mode needs latent 128 */
    mgn_model* m = NULL;
    CHECK(mgn_model_create(&cfg, &m) == MGN_ERR_INVALID && m == NULL);
    CHECK(mgn_last_error(msg, sizeof msg) == MGN_OK && strstr(msg, "128") != NULL);
  }
  /* the parameter table of the reference configuration (fp32 mode needs no device to be built) */
  {
    mgn_model_config cfg = {9, 3, 2, 128, 15, 2, 1e-5f, MGN_COMPUTE_FP32, 0, 0, 0};
    mgn_model* m = NULL;
    int64_t count = 0;
    int32_t n = 0;
    CHECK(mgn_model_create(&cfg, &m) == MGN_OK && m != NULL);
    CHECK(mgn_model_param_count(m, &count) == MGN_OK && count == 2877570);   /* SURVEY 8 a15, L = 4 */
    CHECK(mgn_model_param_layout(m, NULL, 0, &n) == MGN_OK && n == 2 * 10 + 30 * 10 + 8);
    mgn_param_entry first;
    CHECK(mgn_model_param_layout(m, &first, 1, &n) == MGN_OK);
    CHECK(strcmp(first.name, "encoder.node.dense1.weight") == 0 && first.offset == 0 && first.rows == 128 && first.cols == 9);
    CHECK(mgn_model_destroy(m) == MGN_OK);
  }
  /* no CPU fallback: without a device every device entry point reports MGN_ERR_CUDA */
  {
    int32_t ndev = -1;
    CHECK(mgn_device_count(&ndev) == MGN_OK && ndev >= 0);
    if (ndev == 0) {
      mgn_graph* g = NULL;
      const int32_t dummy[2] = {1, 2};
      CHECK(mgn_graph_create(2, 2, dummy, dummy, 1, NULL, &g) == MGN_ERR_CUDA && g == NULL);
      CHECK(mgn_last_error(msg, sizeof msg) == MGN_OK && strlen(msg) > 0);
      float x[4] = {0};
      CHECK(mgn_vec_mul(x, x, 4, x, NULL) == MGN_ERR_CUDA);
      CHECK(mgn_affine_apply(x, 2, 2, 1.f, 0.f, x, 2, 0, NULL) == MGN_ERR_CUDA);
    }
    printf("devices: %d\n", ndev);
  }
  /* multi-GPU transport (SURVEY 8b mgn_dp_*): the symbols link from C; argument checking works without a device or
   * NCCL, and a world of one rank is a no-op that needs neither */
  {
    mgn_comm* comm = NULL;
    unsigned char id[MGN_DP_UNIQUE_ID_BYTES];
    int64_t rows1[1] = {0};
    memset(id, 0, sizeof id);
    CHECK(mgn_dp_unique_id(NULL) == MGN_ERR_INVALID);
    CHECK(mgn_dp_init(id, 2, 2, &comm) == MGN_ERR_INVALID); /* rank outside [0, world) */
    CHECK(comm == NULL);
    CHECK(mgn_dp_init(NULL, 0, 1, &comm) == MGN_ERR_INVALID);
    CHECK(mgn_dp_allreduce(NULL, NULL, 0, MGN_DP_SUM, NULL) == MGN_ERR_INVALID);
    CHECK(mgn_dp_allreduce_normaliser(NULL, NULL, NULL, 0, NULL) == MGN_ERR_INVALID);
    CHECK(mgn_halo_exchange(NULL, NULL, rows1, NULL, rows1, 256, NULL) == MGN_ERR_INVALID);
    CHECK(mgn_dp_rank(NULL, NULL, NULL) == MGN_ERR_INVALID);
    CHECK(mgn_backward_dp(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL, NULL, 4, NULL) == MGN_ERR_INVALID);
    CHECK(mgn_last_error(msg, sizeof msg) == MGN_OK && strlen(msg) > 0);
    CHECK(mgn_dp_finalize(NULL) == MGN_OK);
    CHECK(mgn_library_release() == MGN_OK);
  }
  printf(failures ? "abi_kat: %d check(s) FAILED\n" : "abi_kat: all checks passed\n", failures);
  return failures ? 1 : 0;
}
