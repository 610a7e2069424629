/*
 * CPU restatement of the linear-assignment solver behind the reference's hungarian().
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * /root/reference/utils/hungarian.py:58-65 calls scipy.optimize.linear_sum_assignment, whose algorithm
 * lives in scipy (pinned scipy==1.10.1 in /root/reference/environment.yml:229; scipy/optimize/
 * rectangular_lsap/rectangular_lsap.cpp): the shortest-augmenting-path method of
 * D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016.
 * This file re-states that published algorithm in plain C, keeping the details that decide ties:
 *   - rows are augmented in index order; the matrix is transposed first when it has more rows than columns;
 *   - the `remaining` column list is filled in reverse and shrunk by swap-with-last;
 *   - reduced cost  minVal + cost[i][j] - u[i] - v[j]  is evaluated left to right in double;
 *   - among equal minima an unassigned column wins (the LAST such position scanned), else the FIRST minimum.
 * It is pinned against the live scipy in tests/test_cpu_oracle.py (bit-equal assignments on tie-heavy
 * matrices), and csrc/lap.cu re-states the same traversal for one warp per matrix.
 *
 * lap_ref_solve: cost is row-major float32 [nr x nc] (as the reference hands scipy a float32 array that
 * scipy widens to double).  Writes min(nr, nc) (row, col) pairs sorted by row.  Returns 0, or -1 if
 * infeasible / invalid (NaN or -inf entry), mirroring scipy's ValueError.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int augment(int nc, const double* cost, double* u, double* v, int* path, int* row4col,
                   double* spc, int i, unsigned char* SR, unsigned char* SC, int* remaining,
                   int nr, double* p_minVal) {
  double minVal = 0;
  int num_remaining = nc;
  for (int it = 0; it < nc; it++) remaining[it] = nc - it - 1;
  for (int r = 0; r < nr; r++) SR[r] = 0;
  for (int j = 0; j < nc; j++) { SC[j] = 0; spc[j] = INFINITY; }

  int sink = -1;
  while (sink == -1) {
    int index = -1;
    double lowest = INFINITY;
    SR[i] = 1;
    for (int it = 0; it < num_remaining; it++) {
      int j = remaining[it];
      double r = minVal + cost[(size_t)i * nc + j] - u[i] - v[j];
      if (r < spc[j]) { path[j] = i; spc[j] = r; }
      if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
    }
    minVal = lowest;
    if (minVal == INFINITY) return -1;
    int j = remaining[index];
    if (row4col[j] == -1) sink = j; else i = row4col[j];
    SC[j] = 1;
    remaining[index] = remaining[--num_remaining];
  }
  *p_minVal = minVal;
  return sink;
}

int lap_ref_solve(const float* cost_f32, int nr, int nc, int64_t* rows_out, int64_t* cols_out) {
  if (nr == 0 || nc == 0) return 0;
  const int transpose = nc < nr;
  const int R = transpose ? nc : nr, C = transpose ? nr : nc;
  double* cost = (double*)malloc(sizeof(double) * (size_t)R * C);
  for (int i = 0; i < nr; i++)
    for (int j = 0; j < nc; j++) {
      double x = (double)cost_f32[(size_t)i * nc + j];
      if (x != x || x == -INFINITY) { free(cost); return -1; }
      if (transpose) cost[(size_t)j * C + i] = x; else cost[(size_t)i * C + j] = x;
    }
  double* u = (double*)calloc(R, sizeof(double));
  double* v = (double*)calloc(C, sizeof(double));
  double* spc = (double*)malloc(sizeof(double) * C);
  int* path = (int*)malloc(sizeof(int) * C);
  int* col4row = (int*)malloc(sizeof(int) * R);
  int* row4col = (int*)malloc(sizeof(int) * C);
  int* remaining = (int*)malloc(sizeof(int) * C);
  unsigned char* SR = (unsigned char*)malloc(R);
  unsigned char* SC = (unsigned char*)malloc(C);
  for (int i = 0; i < R; i++) col4row[i] = -1;
  for (int j = 0; j < C; j++) { row4col[j] = -1; path[j] = -1; }
  int rc = 0;
  for (int cur = 0; cur < R; cur++) {
    double minVal;
    int sink = augment(C, cost, u, v, path, row4col, spc, cur, SR, SC, remaining, R, &minVal);
    if (sink < 0) { rc = -1; break; }
    u[cur] += minVal;
    for (int i = 0; i < R; i++)
      if (SR[i] && i != cur) u[i] += minVal - spc[col4row[i]];
    for (int j = 0; j < C; j++)
      if (SC[j]) v[j] -= minVal - spc[j];
    int j = sink;
    while (1) {
      int i = path[j];
      row4col[j] = i;
      int t = col4row[i]; col4row[i] = j; j = t;
      if (i == cur) break;
    }
  }
  if (rc == 0) {
    if (transpose) {
      /* pairs (col4row[i], i) sorted by their first component: walk the original rows in order */
      int k = 0;
      for (int r = 0; r < C; r++)
        if (row4col[r] != -1) { rows_out[k] = r; cols_out[k] = row4col[r]; k++; }
    } else {
      for (int i = 0; i < R; i++) { rows_out[i] = i; cols_out[i] = col4row[i]; }
    }
  }
  free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining);
  free(SR); free(SC);
  return rc;
}
