/* CPU oracle twin in plain C: restatement of GridapHybrid.jl's per-cell static condensation and
 * backward static condensation on the packed batch format.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header): used by tests/ as a checker and by
 * bench.py's cpu_baseline / --impl reference legs as the timed CPU port.  The product library
 * (libgridaphybrid_b200.so) never links or calls it.
 *
 * Follows, per cell (file:line relative to /root/reference):
 *   src/StaticCondensationMap.jl:170-175  densify (zero-fill untouched blocks)
 *   src/StaticCondensationMap.jl:179      dgetrf  -> ora_getrf (LAPACK dgetf2: partial pivoting,
 *                                                   first max |.| as idamax, reciprocal scaling)
 *   src/StaticCondensationMap.jl:183,189  dgetrs  -> ora_getrs (dlaswp + unit-lower + upper solves)
 *   src/StaticCondensationMap.jl:186      dgemm   -> S = A22 - A21*X
 *   src/StaticCondensationMap.jl:192      dgemv   -> g = b2 - A21*y
 *   src/BackwardStaticCondensationMap.jl:91,95,99   dgemv, dgetrf, dgetrs
 * LAPACK/BLAS themselves are third-party (OpenBLAS 0.3.21, Manifest.toml:873-876); the loops below
 * restate the reference (netlib) algorithms and are pinned against SciPy's LAPACK in tests/.
 *
 * Build: see oracle/Makefile (gcc -O3 -pthread -shared -fPIC).  Threads: static slabs of cells over
 * pthreads (libgomp is not in the image).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
  int nfields;
  int n_i, n_b;
  int lenA, lenb;
  const int32_t* ndofs;        /* [nfields] */
  const int64_t* block_offset; /* [nfields*nfields] row-major (i*nfields+j); -1 = untouched */
  const int32_t* perm_fields;  /* [nfields] interior fields then boundary fields, 1-based */
  int n_int_fields;
} ora_plan;

/* dgetf2: unblocked right-looking LU with partial pivoting; col-major, leading dim lda. */
static int ora_getrf(int n, double* a, int lda, int* ipiv) {
  int info = 0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double amax = fabs(a[k + (size_t)k * lda]);
    for (int i = k + 1; i < n; ++i) {
      double v = fabs(a[i + (size_t)k * lda]);
      if (v > amax) { amax = v; p = i; } /* strict '>' keeps the first max (idamax) */
    }
    ipiv[k] = p;
    if (a[p + (size_t)k * lda] != 0.0) {
      if (p != k)
        for (int j = 0; j < n; ++j) {
          double t = a[k + (size_t)j * lda];
          a[k + (size_t)j * lda] = a[p + (size_t)j * lda];
          a[p + (size_t)j * lda] = t;
        }
      double r = 1.0 / a[k + (size_t)k * lda];
      for (int i = k + 1; i < n; ++i) a[i + (size_t)k * lda] *= r;
    } else if (info == 0) {
      info = k + 1;
    }
    for (int j = k + 1; j < n; ++j) {
      double u = a[k + (size_t)j * lda];
      for (int i = k + 1; i < n; ++i) a[i + (size_t)j * lda] -= a[i + (size_t)k * lda] * u;
    }
  }
  return info;
}

/* dgetrs('N'): apply row swaps, solve L (unit) then U; nrhs columns in b (ldb). */
static void ora_getrs(int n, int nrhs, const double* lu, int lda, const int* ipiv, double* b, int ldb) {
  for (int c = 0; c < nrhs; ++c) {
    double* x = b + (size_t)c * ldb;
    for (int k = 0; k < n; ++k) {
      int p = ipiv[k];
      if (p != k) { double t = x[k]; x[k] = x[p]; x[p] = t; }
    }
    for (int k = 0; k < n; ++k) {
      double xk = x[k];
      for (int i = k + 1; i < n; ++i) x[i] -= lu[i + (size_t)k * lda] * xk;
    }
    for (int k = n - 1; k >= 0; --k) {
      x[k] /= lu[k + (size_t)k * lda];
      double xk = x[k];
      for (int i = 0; i < k; ++i) x[i] -= lu[i + (size_t)k * lda] * xk;
    }
  }
}

/* densify the (row fields rf) x (col fields cf) part of one packed cell record into dense col-major */
static void ora_densify(const ora_plan* p, const double* Arec, const int32_t* rf, int nrf, const int32_t* cf,
                        int ncf, double* out, int ld) {
  int co = 0;
  for (int J = 0; J < ncf; ++J) {
    int fj = cf[J] - 1, ro = 0;
    for (int I = 0; I < nrf; ++I) {
      int fi = rf[I] - 1;
      int64_t off = p->block_offset[(size_t)fi * p->nfields + fj];
      int r = p->ndofs[fi], c = p->ndofs[fj];
      for (int jj = 0; jj < c; ++jj)
        for (int ii = 0; ii < r; ++ii)
          out[(ro + ii) + (size_t)(co + jj) * ld] = off < 0 ? 0.0 : Arec[off + ii + (size_t)jj * r];
      ro += r;
    }
    co += p->ndofs[fj];
  }
}

static void ora_densify_vec(const ora_plan* p, const double* brec, const int32_t* F, int nF, double* out) {
  int o = 0;
  for (int I = 0; I < nF; ++I) {
    int f = F[I] - 1, fo = 0;
    for (int q = 0; q < f; ++q) fo += p->ndofs[q];
    for (int l = 0; l < p->ndofs[f]; ++l) out[o++] = brec[fo + l];
  }
}

typedef struct {
  ora_plan p;
  int64_t c0, c1;
  const double *A, *b, *x;
  double *S, *g, *u;
  int32_t* info;
} ora_job;

static void* ora_condense_worker(void* arg) {
  ora_job* j = (ora_job*)arg;
  const ora_plan* p = &j->p;
  const int ni = p->n_i, nb = p->n_b;
  const int32_t* IF = p->perm_fields;
  const int32_t* BF = p->perm_fields + p->n_int_fields;
  const int nif = p->n_int_fields, nbf = p->nfields - p->n_int_fields;
  double* A11 = (double*)malloc(sizeof(double) * (size_t)(ni * ni + 1));
  double* A21 = (double*)malloc(sizeof(double) * (size_t)(nb * ni + 1));
  double* A12 = (double*)malloc(sizeof(double) * (size_t)(ni * nb + 1));
  double* b1 = (double*)malloc(sizeof(double) * (size_t)(ni + 1));
  int* ipiv = (int*)malloc(sizeof(int) * (size_t)(ni + 1));
  for (int64_t c = j->c0; c < j->c1; ++c) {
    const double* Arec = j->A + c * p->lenA;
    const double* brec = j->b + c * p->lenb;
    double* Sc = j->S + c * (int64_t)nb * nb;
    double* gc = j->g + c * (int64_t)nb;
    ora_densify(p, Arec, IF, nif, IF, nif, A11, ni);
    ora_densify(p, Arec, BF, nbf, IF, nif, A21, nb);
    ora_densify(p, Arec, IF, nif, BF, nbf, A12, ni);
    ora_densify(p, Arec, BF, nbf, BF, nbf, Sc, nb);
    ora_densify_vec(p, brec, IF, nif, b1);
    ora_densify_vec(p, brec, BF, nbf, gc);
    int inf = ora_getrf(ni, A11, ni, ipiv);
    j->info[c] = inf;
    if (inf != 0) {
      for (int q = 0; q < nb * nb; ++q) Sc[q] = NAN;
      for (int q = 0; q < nb; ++q) gc[q] = NAN;
      continue;
    }
    ora_getrs(ni, nb, A11, ni, ipiv, A12, ni);
    for (int jj = 0; jj < nb; ++jj) /* dgemm: S -= A21 * X */
      for (int k = 0; k < ni; ++k) {
        double x = A12[k + (size_t)jj * ni];
        for (int i = 0; i < nb; ++i) Sc[i + (size_t)jj * nb] -= A21[i + (size_t)k * nb] * x;
      }
    ora_getrs(ni, 1, A11, ni, ipiv, b1, ni);
    for (int k = 0; k < ni; ++k) /* dgemv: g -= A21 * y */
      for (int i = 0; i < nb; ++i) gc[i] -= A21[i + (size_t)k * nb] * b1[k];
  }
  free(A11); free(A21); free(A12); free(b1); free(ipiv);
  return NULL;
}

static void* ora_backsub_worker(void* arg) {
  ora_job* j = (ora_job*)arg;
  const ora_plan* p = &j->p;
  const int ni = p->n_i, nb = p->n_b;
  const int32_t* IF = p->perm_fields;
  const int32_t* BF = p->perm_fields + p->n_int_fields;
  const int nif = p->n_int_fields, nbf = p->nfields - p->n_int_fields;
  double* A11 = (double*)malloc(sizeof(double) * (size_t)(ni * ni + 1));
  double* A12 = (double*)malloc(sizeof(double) * (size_t)(ni * nb + 1));
  int* ipiv = (int*)malloc(sizeof(int) * (size_t)(ni + 1));
  for (int64_t c = j->c0; c < j->c1; ++c) {
    const double* Arec = j->A + c * p->lenA;
    const double* brec = j->b + c * p->lenb;
    const double* xc = j->x + c * (int64_t)nb;
    double* uc = j->u + c * (int64_t)ni;
    ora_densify(p, Arec, IF, nif, IF, nif, A11, ni);
    ora_densify(p, Arec, IF, nif, BF, nbf, A12, ni);
    ora_densify_vec(p, brec, IF, nif, uc);
    for (int k = 0; k < nb; ++k) /* dgemv: b1 -= A12 * x */
      for (int i = 0; i < ni; ++i) uc[i] -= A12[i + (size_t)k * ni] * xc[k];
    int inf = ora_getrf(ni, A11, ni, ipiv);
    j->info[c] = inf;
    if (inf != 0) { for (int q = 0; q < ni; ++q) uc[q] = NAN; continue; }
    ora_getrs(ni, 1, A11, ni, ipiv, uc, ni);
  }
  free(A11); free(A12); free(ipiv);
  return NULL;
}

int ora_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

static int ora_run(void* (*fn)(void*), ora_job* proto, int64_t ncells, int nthreads) {
  if (nthreads <= 0) nthreads = ora_max_threads();
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > ncells) nthreads = ncells > 0 ? (int)ncells : 1;
  ora_job jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = *proto;
    jobs[t].c0 = ncells * t / nthreads;
    jobs[t].c1 = ncells * (t + 1) / nthreads;
  }
  if (nthreads == 1) { fn(&jobs[0]); return 0; }
  for (int t = 0; t < nthreads; ++t)
    if (pthread_create(&th[t], NULL, fn, &jobs[t]) != 0) return -1;
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}

static void ora_fill_plan(ora_plan* p, int nfields, const int32_t* ndofs, const int64_t* block_offset,
                          const int32_t* perm_fields, int n_int_fields, int64_t lenA, int64_t lenb) {
  p->nfields = nfields; p->n_i = 0; p->n_b = 0; p->lenA = (int)lenA; p->lenb = (int)lenb;
  p->ndofs = ndofs; p->block_offset = block_offset; p->perm_fields = perm_fields; p->n_int_fields = n_int_fields;
  for (int k = 0; k < nfields; ++k) {
    if (k < n_int_fields) p->n_i += ndofs[perm_fields[k] - 1];
    else p->n_b += ndofs[perm_fields[k] - 1];
  }
}

/* Static condensation over a batch.  S [ncells][n_b*n_b] col-major, g [ncells][n_b], info [ncells].
 * nthreads <= 0: all online cores. */
int ora_condense(int nfields, const int32_t* ndofs, const int64_t* block_offset, const int32_t* perm_fields,
                 int n_int_fields, int64_t ncells, const double* A, int64_t lenA, const double* b, int64_t lenb,
                 double* S, double* g, int32_t* info, int nthreads) {
  ora_job j; memset(&j, 0, sizeof(j));
  ora_fill_plan(&j.p, nfields, ndofs, block_offset, perm_fields, n_int_fields, lenA, lenb);
  j.A = A; j.b = b; j.S = S; j.g = g; j.info = info;
  return ora_run(ora_condense_worker, &j, ncells, nthreads);
}

/* Backward static condensation over a batch.  x [ncells][n_b] (cell-wise skeleton values), u [ncells][n_i]. */
int ora_backsub(int nfields, const int32_t* ndofs, const int64_t* block_offset, const int32_t* perm_fields,
                int n_int_fields, int64_t ncells, const double* A, int64_t lenA, const double* b, int64_t lenb,
                const double* x, double* u, int32_t* info, int nthreads) {
  ora_job j; memset(&j, 0, sizeof(j));
  ora_fill_plan(&j.p, nfields, ndofs, block_offset, perm_fields, n_int_fields, lenA, lenb);
  j.A = A; j.b = b; j.x = x; j.u = u; j.info = info;
  return ora_run(ora_backsub_worker, &j, ncells, nthreads);
}

/* ---- assembly of the condensed cells: Gridap's COO numeric loop + SparseArrays.sparse! -------------------------------
 * `assemble_matrix_and_vector(SparseMatrixAssembler(M,L), data)` (/root/reference/src/HybridAffineFEOperators.jl:46) ends
 * in Gridap's numeric loop (cells ascending, `for lj, for li`, push (i,j,v) when both ids are positive) followed
 * by `sparse(I,J,V,m,n)`.  The latter lives in Julia's SparseArrays standard library (not in /root/reference; shipped with
 * the pinned Julia 1.x): its published algorithm `sparse!` is restated here step by step --
 *   (1) row counts -> shifted row pointers, (2) counting sort of (J,V) by row into an unsorted CSR with repeats,
 *   (3) one sweep that combines the repeats of every row in COO order with the single auxiliary array klasttouch and
 *       counts the entries per column, (4) column pointers, (5) counting sort of the CSR into the CSC (rows ascending
 *       inside a column because the rows are visited in order).
 * Serial, like the reference.  Outputs are Julia's: colptr [nfree+1] and rowval 1-based, duplicates summed in COO order,
 * stored zeros kept.  Returns nnz, or -1 if `cap` (capacity of rowval / nzval) is too small, -2 if out of memory.
 * Used by tests (against oracle.julia_sparse) and as the timed CPU baseline of bench.py. */
int64_t ora_assemble_coo_csc(int64_t ncells, int n_b, const int64_t* ids, const double* S, const double* g, int64_t nfree,
                             int64_t* colptr, int64_t* rowval, double* nzval, double* rhs, int64_t cap) {
  const int64_t m = nfree, n = nfree;
  /* Gridap's symbolic pass: count the triplets, then the numeric pass fills them */
  int64_t coolen = 0;
  for (int64_t c = 0; c < ncells; ++c) {
    const int64_t* id = ids + c * n_b;
    int np = 0;
    for (int l = 0; l < n_b; ++l) np += id[l] > 0;
    coolen += (int64_t)np * np;
  }
  int64_t* I = (int64_t*)malloc((size_t)(coolen > 0 ? coolen : 1) * sizeof(int64_t));
  int64_t* J = (int64_t*)malloc((size_t)(coolen > 0 ? coolen : 1) * sizeof(int64_t));
  double* V = (double*)malloc((size_t)(coolen > 0 ? coolen : 1) * sizeof(double));
  int64_t* csrrowptr = (int64_t*)calloc((size_t)m + 2, sizeof(int64_t));
  int64_t* csrcolval = (int64_t*)malloc((size_t)(coolen > 0 ? coolen : 1) * sizeof(int64_t));
  double* csrnzval = (double*)malloc((size_t)(coolen > 0 ? coolen : 1) * sizeof(double));
  int64_t* klasttouch = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  if (!I || !J || !V || !csrrowptr || !csrcolval || !csrnzval || !klasttouch) {
    free(I); free(J); free(V); free(csrrowptr); free(csrcolval); free(csrnzval); free(klasttouch);
    return -2;
  }
  for (int64_t i = 0; i < m; ++i) rhs[i] = 0.0;
  int64_t k = 0;
  for (int64_t c = 0; c < ncells; ++c) {
    const int64_t* id = ids + c * n_b;
    const double* Sc = S + c * (int64_t)n_b * n_b;
    for (int lj = 0; lj < n_b; ++lj) {
      if (id[lj] <= 0) continue;
      for (int li = 0; li < n_b; ++li)
        if (id[li] > 0) { I[k] = id[li]; J[k] = id[lj]; V[k] = Sc[li + (int64_t)lj * n_b]; ++k; }
    }
    for (int li = 0; li < n_b; ++li)
      if (id[li] > 0) rhs[id[li] - 1] += g[c * n_b + li];
  }
  /* sparse!: (1) row counts, shifted forward by one (1-based arithmetic kept: csrrowptr[i] is Julia's csrrowptr[i]) */
  for (k = 0; k < coolen; ++k) csrrowptr[I[k] + 1] += 1;
  int64_t countsum = 1;
  csrrowptr[1] = 1;
  for (int64_t i = 2; i <= m + 1; ++i) { const int64_t ov = csrrowptr[i]; csrrowptr[i] = countsum; countsum += ov; }
  /* (2) counting sort of (J, V) into the CSR arrays; the write positions correct the row pointers */
  for (k = 0; k < coolen; ++k) {
    const int64_t Ik = I[k], csrk = csrrowptr[Ik + 1];
    csrrowptr[Ik + 1] = csrk + 1;
    csrcolval[csrk - 1] = J[k];
    csrnzval[csrk - 1] = V[k];
  }
  /* (3) combine repeats row by row, count per column (colptr used as csccolptr, shifted forward by one) */
  for (int64_t j = 0; j <= n; ++j) colptr[j] = 0;
  int64_t writek = 1, newcsrrowptri = 1, origcsrrowptri = 1, origcsrrowptrip1 = m >= 1 ? csrrowptr[2] : 1;
  for (int64_t i = 1; i <= m; ++i) {
    for (int64_t readk = origcsrrowptri; readk < origcsrrowptrip1; ++readk) {
      const int64_t j = csrcolval[readk - 1];
      if (klasttouch[j] < newcsrrowptri) {
        klasttouch[j] = writek;
        if (writek != readk) { csrcolval[writek - 1] = j; csrnzval[writek - 1] = csrnzval[readk - 1]; }
        ++writek;
        colptr[j] += 1;                                   /* Julia: csccolptr[j+1] += 1 (array is 0-based here) */
      } else {
        const int64_t klt = klasttouch[j];
        csrnzval[klt - 1] = csrnzval[klt - 1] + csrnzval[readk - 1];     /* combine = + , in COO order */
      }
    }
    newcsrrowptri = writek;
    origcsrrowptri = origcsrrowptrip1;
    if (origcsrrowptrip1 != writek) csrrowptr[i + 1] = writek;
    if (i < m) origcsrrowptrip1 = csrrowptr[i + 2];
  }
  /* (4) column pointers: slot j of colptr (0-based) plays Julia's csccolptr[j+1]; after the shift slot j holds the start
   *     of column j (1-based columns), and step (5) advances it to the start of column j+1 -- i.e. slot j ends up as
   *     Julia's final colptr[j+1], slot 0 stays colptr[1] = 1 */
  countsum = 1;
  colptr[0] = 1;
  for (int64_t j = 1; j <= n; ++j) { const int64_t ov = colptr[j]; colptr[j] = countsum; countsum += ov; }
  const int64_t nnz = countsum - 1;
  int64_t rc = nnz;
  if (nnz > cap) {
    rc = -1;
  } else {
    /* (5) counting sort of the CSR into the CSC; the write positions correct the column pointers */
    for (int64_t i = 1; i <= m; ++i) {
      for (int64_t csrk = csrrowptr[i]; csrk < csrrowptr[i + 1]; ++csrk) {
        const int64_t j = csrcolval[csrk - 1];
        const int64_t csck = colptr[j];
        colptr[j] = csck + 1;
        rowval[csck - 1] = i;
        nzval[csck - 1] = csrnzval[csrk - 1];
      }
    }
  }
  free(I); free(J); free(V); free(csrrowptr); free(csrcolval); free(csrnzval); free(klasttouch);
  return rc;
}
