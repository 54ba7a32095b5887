/* c_abi_demo.c — libmrscore.so driven from plain C99 through include/mrscore.h only: the same sequence a JNI / JNA / Panama binding
 * issues.  Build:  gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lmusicrecommendation_b200 -lmrscore -o c_abi_demo
 * The data set is the hand-derived fixture of SURVEY.md §4.3 (3 train users, 2 test users, 4 songs).  Without a CUDA device mr_create
 * reports MR_ERR_CUDA (there is no CPU fallback) and the program exits with status 3. */
#include <stdio.h>
#include "mrscore.h"

int main(void) {
  /* train: A = {s1, s2}, B = {s2, s3}, C = {s3};  test (visible half): X = {s1, s4}, Y = {s2} */
  const int64_t tr_ptr[] = {0, 2, 4, 5};
  const int32_t tr_col[] = {0, 1, 1, 2, 2};
  const int64_t te_ptr[] = {0, 2, 3};
  const int32_t te_col[] = {0, 3, 1};
  const int32_t deg_train[] = {2, 2, 1}, deg_test[] = {2, 1};
  const int32_t deg_song_all[] = {2, 3, 2, 1};          /* train + test-visible listeners (MusicRecommender.scala:41, 53, 237) */
  mr_handle* h = NULL;
  const int device = 0;
  int rc = mr_create(&h, &device, 1, MR_ENGINE_AUTO);
  if (rc != MR_OK) {
    fprintf(stderr, "mr_create: rc=%d (%s)\n", rc, mr_last_error(h));
    mr_destroy(h);
    return rc;
  }
  rc = mr_load(h, 3, 2, 4, tr_ptr, tr_col, te_ptr, te_col, deg_train, deg_test, deg_song_all);
  if (rc == MR_OK) {
    int32_t song[2 * 2], len[2];
    double score[2 * 2];
    rc = mr_topk(h, MR_UBM, 0.0, 0, 2, song, score, len);
    if (rc == MR_OK)
      for (int u = 0; u < 2; ++u)
        for (int i = 0; i < len[u]; ++i) printf("user %d  #%d  song %d  score %.17g\n", u, i + 1, song[u * 2 + i], score[u * 2 + i]);
  }
  if (rc != MR_OK) fprintf(stderr, "rc=%d (%s)\n", rc, mr_last_error(h));
  mr_destroy(h);
  return rc;
}
