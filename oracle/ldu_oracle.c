/*
 * ldu_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).  See ldu_oracle.h for
 * scope, the reference files restated and the "parity unpinned" statement.
 *
 * Compile: gcc -O3 -ffp-contract=off -fPIC -shared (see Makefile).
 */
#include "ldu_oracle.h"

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_GREAT 1.0e+20  /* lduMatrix::great_ */
#define ORC_SMALL 1.0e-20  /* lduMatrix::small_ */
#define ORC_VSMALL 1.0e-300 /* VSMALL (checkSingularity) */
#define ORC_SMALL_ 1.0e-15 /* SMALL (relTol > SMALL test in checkConvergence) */

typedef struct orc_iface
{
    int kind;
    int nFaces;
    int* faceCells;
    double* bouCoeffs;
    double* intCoeffs;
    int peerRow, peerIface;
    int* ggiOffsets; /* NULL => identity */
    int* ggiAddr;
    double* ggiWeights;
    double* buf; /* matrixUpdateBuffer_ written by init of THIS patch; size = peer nFaces */
    int bufSize;
    /* shadow patch spread over several rows (decomposed regionCouple pair, interpolated on the global zones): zone face z
       of the shadow is face zonePos[z] of interface zoneIface[z] of row zoneRow[z]; pnf = this patch's own neighbour values */
    int nZone;
    int *zoneRow, *zoneIface, *zonePos;
    double* pnf;
} orc_iface;

typedef struct orc_row
{
    int rank, region;
    int nCells, nFaces;
    int offset; /* into concatenated vectors */
    int *l, *u;
    int* losort; /* faces sorted by u (stable): lduAddressing::losortAddr */
    double *diag, *upper, *lower; /* lower == upper when symmetric */
    int symmetric;
    int nIfaces, capIfaces;
    orc_iface* ifaces;
    double* rD;
    int precond; /* effective per-row type after Cholesky resolution */
} orc_row;

struct orc_sys
{
    int nRows, nRanks;
    orc_row* rows;
    int total;
    int precond;
    double* rankPartial;
    double* rowPartial; /* [nRows] scratch of the threaded reductions */
    int redMode; /* 0: sequential sums (the reference); 1: pairwise sums -- only to MEASURE how sensitive a
                    residual history is to the summation order (tests), never the parity target */
};

const char* orc_version(void) { return "ldu_oracle 1.0 (foam-extend-4.1 restatement, parity unpinned)"; }

/* ------------------------------------------------------------------ build */

orc_sys* orc_create(int nRows, int nRanks)
{
    orc_sys* s = (orc_sys*)calloc(1, sizeof(orc_sys));
    s->nRows = nRows;
    s->nRanks = nRanks > 0 ? nRanks : 1;
    s->rows = (orc_row*)calloc((size_t)nRows, sizeof(orc_row));
    s->rankPartial = (double*)calloc((size_t)s->nRanks, sizeof(double));
    s->rowPartial = (double*)calloc((size_t)(s->nRows > 0 ? s->nRows : 1), sizeof(double));
    s->precond = -1;
    return s;
}

void orc_destroy(orc_sys* s)
{
    if (!s) return;
    for (int r = 0; r < s->nRows; r++)
    {
        orc_row* R = &s->rows[r];
        free(R->l);
        free(R->u);
        free(R->losort);
        free(R->diag);
        if (R->lower != R->upper) free(R->lower);
        free(R->upper);
        free(R->rD);
        for (int i = 0; i < R->nIfaces; i++)
        {
            orc_iface* I = &R->ifaces[i];
            free(I->faceCells);
            free(I->bouCoeffs);
            free(I->intCoeffs);
            free(I->ggiOffsets);
            free(I->ggiAddr);
            free(I->ggiWeights);
            free(I->buf);
            free(I->zoneRow);
            free(I->zoneIface);
            free(I->zonePos);
            free(I->pnf);
        }
        free(R->ifaces);
    }
    free(s->rows);
    free(s->rankPartial);
    free(s->rowPartial);
    free(s);
}

static void* dupmem(const void* p, size_t n)
{
    void* q = malloc(n ? n : 1);
    if (p && n) memcpy(q, p, n);
    return q;
}

static void recompute_offsets(orc_sys* s)
{
    int off = 0;
    for (int r = 0; r < s->nRows; r++)
    {
        s->rows[r].offset = off;
        off += s->rows[r].nCells;
    }
    s->total = off;
}

int orc_set_row(orc_sys* s, int row, int rank, int region, int nCells, int nFaces,
                const int* lowerAddr, const int* upperAddr)
{
    if (row < 0 || row >= s->nRows || rank < 0 || rank >= s->nRanks) return -1;
    orc_row* R = &s->rows[row];
    R->rank = rank;
    R->region = region;
    R->nCells = nCells;
    R->nFaces = nFaces;
    R->l = (int*)dupmem(lowerAddr, sizeof(int) * (size_t)nFaces);
    R->u = (int*)dupmem(upperAddr, sizeof(int) * (size_t)nFaces);
    for (int f = 0; f < nFaces; f++)
        if (R->l[f] < 0 || R->u[f] >= nCells || R->l[f] >= R->u[f]) return -2;
    /* losort: stable counting sort of faces by upper address (lduAddressing::calcLosort) */
    R->losort = (int*)malloc(sizeof(int) * (size_t)(nFaces ? nFaces : 1));
    int* cnt = (int*)calloc((size_t)nCells + 1, sizeof(int));
    for (int f = 0; f < nFaces; f++) cnt[R->u[f] + 1]++;
    for (int c = 0; c < nCells; c++) cnt[c + 1] += cnt[c];
    for (int f = 0; f < nFaces; f++) R->losort[cnt[R->u[f]]++] = f;
    free(cnt);
    recompute_offsets(s);
    return 0;
}

int orc_set_coeffs(orc_sys* s, int row, const double* diag, const double* upper, const double* lower)
{
    if (row < 0 || row >= s->nRows) return -1;
    orc_row* R = &s->rows[row];
    free(R->diag);
    if (R->lower != R->upper) free(R->lower);
    free(R->upper);
    R->diag = (double*)dupmem(diag, sizeof(double) * (size_t)R->nCells);
    R->upper = (double*)dupmem(upper, sizeof(double) * (size_t)R->nFaces);
    if (lower)
    {
        R->lower = (double*)dupmem(lower, sizeof(double) * (size_t)R->nFaces);
        R->symmetric = 0;
    }
    else
    {
        R->lower = R->upper;
        R->symmetric = 1;
    }
    s->precond = -1;
    return 0;
}

int orc_add_iface(orc_sys* s, int row, int kind, int nFaces, const int* faceCells,
                  const double* bouCoeffs, const double* intCoeffs, int peerRow, int peerIface,
                  const int* ggiOffsets, const int* ggiAddr, const double* ggiWeights)
{
    if (row < 0 || row >= s->nRows) return -1;
    orc_row* R = &s->rows[row];
    if (R->nIfaces == R->capIfaces)
    {
        R->capIfaces = R->capIfaces ? 2 * R->capIfaces : 4;
        R->ifaces = (orc_iface*)realloc(R->ifaces, sizeof(orc_iface) * (size_t)R->capIfaces);
    }
    orc_iface* I = &R->ifaces[R->nIfaces];
    memset(I, 0, sizeof(*I));
    I->kind = kind;
    I->nFaces = nFaces;
    I->faceCells = (int*)dupmem(faceCells, sizeof(int) * (size_t)nFaces);
    I->bouCoeffs = (double*)dupmem(bouCoeffs, sizeof(double) * (size_t)nFaces);
    /* NULL internal coefficients: the boundary coefficients are used for the transposed product too
       (b200_ldu.h, b200_sys_set_interface_coeffs; same rule in the device library) */
    I->intCoeffs = (double*)dupmem(intCoeffs ? intCoeffs : bouCoeffs, sizeof(double) * (size_t)(nFaces ? nFaces : 1));
    I->peerRow = peerRow;
    I->peerIface = peerIface;
    if (ggiOffsets)
    {
        int nnz = ggiOffsets[nFaces];
        I->ggiOffsets = (int*)dupmem(ggiOffsets, sizeof(int) * ((size_t)nFaces + 1));
        I->ggiAddr = (int*)dupmem(ggiAddr, sizeof(int) * (size_t)nnz);
        I->ggiWeights = (double*)dupmem(ggiWeights, sizeof(double) * (size_t)nnz);
    }
    for (int i = 0; i < nFaces; i++)
        if (faceCells[i] < 0 || faceCells[i] >= R->nCells) return -2;
    return R->nIfaces++;
}

/* Decomposed regionCouple pair (regionCouplePolyPatch not localParallel()): the GGI addresses of interface (row, iface) are
 * face labels of the shadow's global ZONE; zoneRow/zoneIface/zonePos say where each zone face lives (-1: not held by anybody,
 * must not be addressed).  foam-extend expands the shadow's patchInternalField to the zone, interpolates with the zone
 * addressing and filters to the local faces: pnf[i] = sum_k zoneField[addr[k]] * w[k], evaluated here on the receiving side. */
int orc_set_iface_zone(orc_sys* s, int row, int iface, int nZone, const int* zoneRow, const int* zoneIface, const int* zonePos)
{
    if (row < 0 || row >= s->nRows) return -1;
    orc_row* R = &s->rows[row];
    if (iface < 0 || iface >= R->nIfaces) return -1;
    orc_iface* I = &R->ifaces[iface];
    if (!I->ggiOffsets || nZone < 0) return -2;
    for (int k = 0; k < I->ggiOffsets[I->nFaces]; k++)
        if (I->ggiAddr[k] < 0 || I->ggiAddr[k] >= nZone || zoneRow[I->ggiAddr[k]] < 0) return -3;
    I->nZone = nZone;
    I->zoneRow = (int*)dupmem(zoneRow, sizeof(int) * (size_t)(nZone ? nZone : 1));
    I->zoneIface = (int*)dupmem(zoneIface, sizeof(int) * (size_t)(nZone ? nZone : 1));
    I->zonePos = (int*)dupmem(zonePos, sizeof(int) * (size_t)(nZone ? nZone : 1));
    I->pnf = (double*)malloc(sizeof(double) * (size_t)(I->nFaces ? I->nFaces : 1));
    return 0;
}

int orc_total_cells(const orc_sys* s) { return s->total; }

/* ------------------------------------------------------------ reductions */
/* kind: 0 sum a*b, 1 sum |a|, 2 sum a, 3 sum |a-c| + |b-c| */
static double red_term(int kind, const double* a, const double* b, const double* c, int i)
{
    switch (kind)
    {
        case 0: return a[i] * b[i];
        case 1: return fabs(a[i]);
        case 2: return a[i];
        default: return fabs(a[i] - c[i]) + fabs(b[i] - c[i]);
    }
}
static double red_pairwise(int kind, const double* a, const double* b, const double* c, int lo, int hi)
{
    if (hi - lo <= 64)
    {
        double s = 0.0;
        for (int i = lo; i < hi; i++) s += red_term(kind, a, b, c, i);
        return s;
    }
    int mid = lo + (hi - lo) / 2;
    return red_pairwise(kind, a, b, c, lo, mid) + red_pairwise(kind, a, b, c, mid, hi);
}
static double red_rows(const orc_sys* s, int kind, const double* a, const double* b, const double* c)
{
    for (int k = 0; k < s->nRanks; k++) s->rankPartial[k] = 0.0;
    for (int r = 0; r < s->nRows; r++)
    {
        const orc_row* R = &s->rows[r];
        const double* pa = a + R->offset;
        const double* pb = b ? b + R->offset : NULL;
        const double* pc = c ? c + R->offset : NULL;
        s->rankPartial[R->rank] += red_pairwise(kind, pa, pb, pc, 0, R->nCells);
    }
    double tot = s->rankPartial[0];
    for (int k = 1; k < s->nRanks; k++) tot += s->rankPartial[k];
    return tot;
}
void orc_set_reduction_mode(orc_sys* s, int mode) { s->redMode = mode; }

/* Threads stand in for the MPI ranks of a decomposed foam-extend run: the loops over the
 * rows (= sub-domain matrices) and the element-wise vector updates are shared out; every sum keeps its order
 * (per-row partial sums, combined in row / rank order), so the results do not depend on the thread count. */
static int g_threads = 1;
typedef void (*orc_job)(int i, void* ctx);
static struct
{
    pthread_t th[64];
    int nWorkers;                 /* started worker threads (the caller works too) */
    pthread_mutex_t mu;
    pthread_cond_t cv;
    unsigned long gen;            /* job generation */
    orc_job fn;
    void* ctx;
    int n;
    volatile int next, finished;  /* next index to hand out; workers that left the current job */
} g_pool = {.nWorkers = 0, .mu = PTHREAD_MUTEX_INITIALIZER, .cv = PTHREAD_COND_INITIALIZER};

static void pool_run_items(void)
{
    for (;;)
    {
        int i = __atomic_fetch_add(&g_pool.next, 1, __ATOMIC_RELAXED);
        if (i >= g_pool.n) break;
        g_pool.fn(i, g_pool.ctx);
    }
}
static void* pool_worker(void* arg)
{
    /* a worker starts at the generation that was current when it was created (par_for creates workers before it
       publishes a job), so it never takes part in - or reports completion of - a job older than itself */
    unsigned long seen = (unsigned long)(size_t)arg;
    for (;;)
    {
        pthread_mutex_lock(&g_pool.mu);
        while (g_pool.gen == seen) pthread_cond_wait(&g_pool.cv, &g_pool.mu);
        seen = g_pool.gen;
        pthread_mutex_unlock(&g_pool.mu);
        pool_run_items();
        __atomic_fetch_add(&g_pool.finished, 1, __ATOMIC_RELEASE);
    }
    return NULL;
}
/* f(i, ctx) for i in [0, n): items are handed out dynamically; the caller takes part */
static void par_for(int n, orc_job f, void* ctx)
{
    if (g_threads <= 1 || n <= 1)
    {
        for (int i = 0; i < n; i++) f(i, ctx);
        return;
    }
    while (g_pool.nWorkers < g_threads - 1 && g_pool.nWorkers < 64)
    {
        if (pthread_create(&g_pool.th[g_pool.nWorkers], NULL, pool_worker, (void*)(size_t)g_pool.gen)) break;
        g_pool.nWorkers++;
    }
    pthread_mutex_lock(&g_pool.mu);
    g_pool.fn = f;
    g_pool.ctx = ctx;
    g_pool.n = n;
    g_pool.next = 0;
    g_pool.finished = 0;
    g_pool.gen++;
    pthread_cond_broadcast(&g_pool.cv);
    pthread_mutex_unlock(&g_pool.mu);
    pool_run_items();
    while (__atomic_load_n(&g_pool.finished, __ATOMIC_ACQUIRE) < g_pool.nWorkers) sched_yield();
}
void orc_set_threads(int n) { g_threads = n > 0 ? (n > 65 ? 65 : n) : 1; }
int orc_get_threads(void) { return g_threads; }

/* element-wise vector updates: chunk i of nChunks */
typedef struct
{
    int kind, n, nChunks;
    double *y, *y2;
    const double *a, *b, *c;
    double s1, s2;
} vec_job;
static void vec_chunk(int ch, void* vctx)
{
    vec_job* J = (vec_job*)vctx;
    const int lo = (int)((long long)J->n * ch / J->nChunks), hi = (int)((long long)J->n * (ch + 1) / J->nChunks);
    double* y = J->y;
    const double *a = J->a, *b = J->b, *c = J->c;
    const double s1 = J->s1, s2 = J->s2;
    switch (J->kind)
    {
        case 0: for (int i = lo; i < hi; i++) y[i] = a[i] + s1 * y[i] - s1 * s2 * b[i]; break; /* p = r + beta p - beta omega v */
        case 1: for (int i = lo; i < hi; i++) y[i] = a[i] - s1 * b[i]; break;                    /* s = r - alpha v */
        case 2: for (int i = lo; i < hi; i++) y[i] = y[i] + s1 * a[i] + s2 * b[i]; break;        /* x += alpha ph + omega sh */
        default: for (int i = lo; i < hi; i++) y[i] = a[i] - s1 * b[i]; break;
    }
    (void)c;
}
static void vec_op(int kind, int n, double* y, const double* a, const double* b, double s1, double s2)
{
    vec_job J = {.kind = kind, .n = n, .nChunks = g_threads > 1 ? 4 * g_threads : 1, .y = y, .a = a, .b = b, .s1 = s1, .s2 = s2};
    par_for(J.nChunks, vec_chunk, &J);
}

/* gSumProd / gSumMag / gSum over a FieldField: per rank, rows in list order,
 * cells sequentially (FieldFunctions.C sumProd / FieldFieldFunctions.C), then
 * reduce(sum) over ranks (taken in ascending rank order). */

typedef struct
{
    const orc_sys* s;
    const double *a, *b;
} red_job;
static void row_sumprod(int r, void* ctx)
{
    const red_job* J = (const red_job*)ctx;
    const orc_row* R = &J->s->rows[r];
    const double* pa = J->a + R->offset;
    const double* pb = J->b + R->offset;
    double sum = 0.0;
    for (int c = 0; c < R->nCells; c++) sum += pa[c] * pb[c];
    J->s->rowPartial[r] = sum;
}
static void row_summag(int r, void* ctx)
{
    const red_job* J = (const red_job*)ctx;
    const orc_row* R = &J->s->rows[r];
    const double* pa = J->a + R->offset;
    double sum = 0.0;
    for (int c = 0; c < R->nCells; c++) sum += fabs(pa[c]);
    J->s->rowPartial[r] = sum;
}

double orc_gsumprod(const orc_sys* s, const double* a, const double* b)
{
    if (s->redMode) return red_rows(s, 0, a, b, NULL);
    for (int k = 0; k < s->nRanks; k++) s->rankPartial[k] = 0.0;
    red_job J = {s, a, b};
    par_for(s->nRows, row_sumprod, &J);
    for (int r = 0; r < s->nRows; r++) s->rankPartial[s->rows[r].rank] += s->rowPartial[r];
    double tot = s->rankPartial[0];
    for (int k = 1; k < s->nRanks; k++) tot += s->rankPartial[k];
    return tot;
}

double orc_gsummag(const orc_sys* s, const double* a)
{
    if (s->redMode) return red_rows(s, 1, a, NULL, NULL);
    for (int k = 0; k < s->nRanks; k++) s->rankPartial[k] = 0.0;
    red_job J = {s, a, NULL};
    par_for(s->nRows, row_summag, &J);
    for (int r = 0; r < s->nRows; r++) s->rankPartial[s->rows[r].rank] += s->rowPartial[r];
    double tot = s->rankPartial[0];
    for (int k = 1; k < s->nRanks; k++) tot += s->rankPartial[k];
    return tot;
}

static double gsum(const orc_sys* s, const double* a)
{
    if (s->redMode) return red_rows(s, 2, a, NULL, NULL);
    for (int k = 0; k < s->nRanks; k++) s->rankPartial[k] = 0.0;
    for (int r = 0; r < s->nRows; r++)
    {
        const orc_row* R = &s->rows[r];
        const double* pa = a + R->offset;
        double sum = 0.0;
        for (int c = 0; c < R->nCells; c++) sum += pa[c];
        s->rankPartial[R->rank] += sum;
    }
    double tot = s->rankPartial[0];
    for (int k = 1; k < s->nRanks; k++) tot += s->rankPartial[k];
    return tot;
}

/* gSum(mag(Ax - tmp) + mag(b - tmp)) */
static double gsum_normterms(const orc_sys* s, const double* Ax, const double* b, const double* tmp)
{
    if (s->redMode) return red_rows(s, 3, Ax, b, tmp);
    for (int k = 0; k < s->nRanks; k++) s->rankPartial[k] = 0.0;
    for (int r = 0; r < s->nRows; r++)
    {
        const orc_row* R = &s->rows[r];
        int o = R->offset;
        double sum = 0.0;
        for (int c = 0; c < R->nCells; c++)
            sum += fabs(Ax[o + c] - tmp[o + c]) + fabs(b[o + c] - tmp[o + c]);
        s->rankPartial[R->rank] += sum;
    }
    double tot = s->rankPartial[0];
    for (int k = 1; k < s->nRanks; k++) tot += s->rankPartial[k];
    return tot;
}

/* ------------------------------------------------------------ interfaces */

void orc_ggi_interpolate(int nTo, const int* offsets, const int* addr, const double* weights,
                         const double* ff, int nComp, double* result)
{
    /* GGIInterpolate.C: result zero-initialised; result[faceI] += ff[curAddr[i]]*curWeights[i] */
    for (int i = 0; i < nTo; i++)
        for (int d = 0; d < nComp; d++)
        {
            double acc = 0.0;
            for (int k = offsets[i]; k < offsets[i + 1]; k++) acc += ff[addr[k] * nComp + d] * weights[k];
            result[i * nComp + d] = acc;
        }
}

/* init phase of one interface: fill buf of THIS patch with this side's
 * patchInternalField mapped onto the PEER's faces
 * (monolithicCouplingFvPatchField.C:392-405; processorFvPatchField send). */
static int iface_init(orc_sys* s, orc_row* R, orc_iface* I, const double* x)
{
    if (I->zoneRow)
    { /* shadow spread over rows: gather the zone values this patch's rows address, in addressing order */
        for (int i = 0; i < I->nFaces; i++)
        {
            double acc = 0.0;
            for (int k = I->ggiOffsets[i]; k < I->ggiOffsets[i + 1]; k++)
            {
                const int z = I->ggiAddr[k];
                const orc_row* SR = &s->rows[I->zoneRow[z]];
                if (I->zoneIface[z] < 0 || I->zoneIface[z] >= SR->nIfaces) return -3;
                const orc_iface* Q = &SR->ifaces[I->zoneIface[z]];
                if (I->zonePos[z] < 0 || I->zonePos[z] >= Q->nFaces) return -3;
                acc += x[SR->offset + Q->faceCells[I->zonePos[z]]] * I->ggiWeights[k];
            }
            I->pnf[i] = acc;
        }
        return 0;
    }
    if (I->peerRow < 0 || I->peerRow >= s->nRows) return -3;
    orc_row* PR = &s->rows[I->peerRow];
    if (I->peerIface < 0 || I->peerIface >= PR->nIfaces) return -3;
    orc_iface* Q = &PR->ifaces[I->peerIface];
    if (Q->zoneRow) return 0; /* the peer gathers its neighbour values from the zone itself */
    const double* xr = x + R->offset;
    int nOut = Q->nFaces;
    if (I->bufSize < nOut)
    {
        free(I->buf);
        I->buf = (double*)malloc(sizeof(double) * (size_t)(nOut ? nOut : 1));
        I->bufSize = nOut;
    }
    if (Q->ggiOffsets)
    {
        for (int i = 0; i < nOut; i++)
        {
            double acc = 0.0;
            for (int k = Q->ggiOffsets[i]; k < Q->ggiOffsets[i + 1]; k++)
                acc += xr[I->faceCells[Q->ggiAddr[k]]] * Q->ggiWeights[k];
            I->buf[i] = acc;
        }
    }
    else
    {
        if (nOut != I->nFaces) return -4;
        for (int i = 0; i < nOut; i++) I->buf[i] = xr[I->faceCells[i]];
    }
    return 0;
}

/* update phase: pnf = shadow().matrixUpdateBuffer(); result[fc[i]] -= coeffs[i]*pnf[i]
 * (+= when switchToLhs)  -- monolithicCouplingFvPatchField.C:430-455 */
static void iface_update(orc_sys* s, orc_row* R, orc_iface* I, const double* coeffs, double* y,
                         int switchToLhs)
{
    const double* pnf = I->zoneRow ? I->pnf : s->rows[I->peerRow].ifaces[I->peerIface].buf;
    double* yr = y + R->offset;
    if (switchToLhs)
        for (int i = 0; i < I->nFaces; i++) yr[I->faceCells[i]] += coeffs[i] * pnf[i];
    else
        for (int i = 0; i < I->nFaces; i++) yr[I->faceCells[i]] -= coeffs[i] * pnf[i];
}

/* coupledLduMatrix::initMatrixInterfaces: all non-processor interfaces of all
 * rows first, then all processor interfaces. */
static int init_interfaces(orc_sys* s, const double* x)
{
    for (int phase = 0; phase < 2; phase++)
        for (int r = 0; r < s->nRows; r++)
        {
            orc_row* R = &s->rows[r];
            for (int i = 0; i < R->nIfaces; i++)
            {
                orc_iface* I = &R->ifaces[i];
                if ((I->kind == ORC_IFACE_PROCESSOR) != (phase == 1)) continue;
                int rc = iface_init(s, R, I, x);
                if (rc) return rc;
            }
        }
    return 0;
}

static void update_interfaces(orc_sys* s, double* y, int useInt, int switchToLhs)
{
    for (int phase = 0; phase < 2; phase++)
        for (int r = 0; r < s->nRows; r++)
        {
            orc_row* R = &s->rows[r];
            for (int i = 0; i < R->nIfaces; i++)
            {
                orc_iface* I = &R->ifaces[i];
                if ((I->kind == ORC_IFACE_PROCESSOR) != (phase == 1)) continue;
                iface_update(s, R, I, useInt ? I->intCoeffs : I->bouCoeffs, y, switchToLhs);
            }
        }
}

/* --------------------------------------------------------------- products */

typedef struct
{
    orc_sys* s;
    const double* x;
    double* y;
} mul_job;
static void row_amul(int r, void* ctx)
{
    /* lduMatrix::AmulCore (lduMatrixATmul.C) */
    const mul_job* J = (const mul_job*)ctx;
    const orc_row* R = &J->s->rows[r];
    const double* psi = J->x + R->offset;
    double* Apsi = J->y + R->offset;
    for (int c = 0; c < R->nCells; c++) Apsi[c] = R->diag[c] * psi[c];
    for (int f = 0; f < R->nFaces; f++)
    {
        Apsi[R->u[f]] += R->lower[f] * psi[R->l[f]];
        Apsi[R->l[f]] += R->upper[f] * psi[R->u[f]];
    }
}
static void row_tmul(int r, void* ctx)
{
    /* lduMatrix::TmulCore */
    const mul_job* J = (const mul_job*)ctx;
    const orc_row* R = &J->s->rows[r];
    const double* psi = J->x + R->offset;
    double* Tpsi = J->y + R->offset;
    for (int c = 0; c < R->nCells; c++) Tpsi[c] = R->diag[c] * psi[c];
    for (int f = 0; f < R->nFaces; f++)
    {
        Tpsi[R->u[f]] += R->upper[f] * psi[R->l[f]];
        Tpsi[R->l[f]] += R->lower[f] * psi[R->u[f]];
    }
}

int orc_amul(orc_sys* s, const double* x, double* y)
{
    for (int i = 0; i < s->total; i++) y[i] = 0.0; /* coupledLduMatrix::Amul: result = 0 */
    int rc = init_interfaces(s, x);
    if (rc) return rc;
    mul_job J = {s, x, y};
    par_for(s->nRows, row_amul, &J);
    update_interfaces(s, y, 0, 0);
    return 0;
}

int orc_tmul(orc_sys* s, const double* x, double* y)
{
    for (int i = 0; i < s->total; i++) y[i] = 0.0;
    int rc = init_interfaces(s, x);
    if (rc) return rc;
    mul_job J = {s, x, y};
    par_for(s->nRows, row_tmul, &J);
    update_interfaces(s, y, 1, 0);
    return 0;
}

int orc_sumA(orc_sys* s, double* sumA)
{
    for (int r = 0; r < s->nRows; r++)
    {
        const orc_row* R = &s->rows[r];
        double* sa = sumA + R->offset;
        for (int c = 0; c < R->nCells; c++) sa[c] = R->diag[c];
        for (int f = 0; f < R->nFaces; f++)
        {
            sa[R->u[f]] += R->lower[f];
            sa[R->l[f]] += R->upper[f];
        }
        for (int i = 0; i < R->nIfaces; i++)
        {
            const orc_iface* I = &R->ifaces[i];
            for (int k = 0; k < I->nFaces; k++) sa[I->faceCells[k]] -= I->bouCoeffs[k];
        }
    }
    return 0;
}

int orc_residual(orc_sys* s, const double* x, const double* b, double* res)
{
    int rc = init_interfaces(s, x);
    if (rc) return rc;
    for (int r = 0; r < s->nRows; r++)
    {
        const orc_row* R = &s->rows[r];
        const double* psi = x + R->offset;
        const double* src = b + R->offset;
        double* rA = res + R->offset;
        for (int c = 0; c < R->nCells; c++) rA[c] = src[c] - R->diag[c] * psi[c];
        for (int f = 0; f < R->nFaces; f++)
        {
            rA[R->u[f]] -= R->lower[f] * psi[R->l[f]];
            rA[R->l[f]] -= R->upper[f] * psi[R->u[f]];
        }
    }
    update_interfaces(s, res, 0, 1);
    return 0;
}

/* --------------------------------------------------------- preconditioners */

int orc_precond_setup(orc_sys* s, int precond)
{
    if (precond < ORC_PRECOND_NONE || precond > ORC_PRECOND_CHOLESKY) return -1;
    for (int r = 0; r < s->nRows; r++)
    {
        orc_row* R = &s->rows[r];
        if (!R->diag) return -2;
        free(R->rD);
        R->rD = (double*)malloc(sizeof(double) * (size_t)(R->nCells ? R->nCells : 1));
        int eff = precond;
        if (precond == ORC_PRECOND_CHOLESKY) eff = R->symmetric ? ORC_PRECOND_DIC : ORC_PRECOND_DILU;
        R->precond = eff;
        double* rD = R->rD;
        for (int c = 0; c < R->nCells; c++) rD[c] = R->diag[c];
        if (eff == ORC_PRECOND_DIC)
        {
            /* DICPreconditioner::calcReciprocalD / CholeskyPrecon::calcPreconDiag (symmetric) */
            for (int f = 0; f < R->nFaces; f++)
                rD[R->u[f]] -= R->upper[f] * R->upper[f] / rD[R->l[f]];
        }
        else if (eff == ORC_PRECOND_DILU)
        {
            /* DILUPreconditioner::calcReciprocalD / CholeskyPrecon (asymmetric) */
            for (int f = 0; f < R->nFaces; f++)
                rD[R->u[f]] -= R->upper[f] * R->lower[f] / rD[R->l[f]];
        }
        if (eff != ORC_PRECOND_NONE)
            for (int c = 0; c < R->nCells; c++) rD[c] = 1.0 / rD[c];
    }
    s->precond = precond;
    return 0;
}

int orc_get_rD(orc_sys* s, double* out)
{
    if (s->precond < 0) return -1;
    for (int r = 0; r < s->nRows; r++)
    {
        const orc_row* R = &s->rows[r];
        memcpy(out + R->offset, R->rD, sizeof(double) * (size_t)R->nCells);
    }
    return 0;
}

typedef struct
{
    orc_sys* s;
    const double* rIn;
    double* wOut;
    int transpose;
} pre_job;
static void row_precondition(int r, void* ctx)
{
    const pre_job* J = (const pre_job*)ctx;
    const orc_row* R = &J->s->rows[r];
    const double* rA = J->rIn + R->offset;
    double* wA = J->wOut + R->offset;
    const double* rD = R->rD;
    const int nF = R->nFaces;
    if (R->precond == ORC_PRECOND_NONE)
    {
        for (int c = 0; c < R->nCells; c++) wA[c] = rA[c];
        return;
    }
    for (int c = 0; c < R->nCells; c++) wA[c] = rD[c] * rA[c];
    if (R->precond == ORC_PRECOND_DIC)
    {
        for (int f = 0; f < nF; f++) wA[R->u[f]] -= rD[R->u[f]] * R->upper[f] * wA[R->l[f]];
        for (int f = nF - 1; f >= 0; f--) wA[R->l[f]] -= rD[R->l[f]] * R->upper[f] * wA[R->u[f]];
    }
    else if (R->precond == ORC_PRECOND_DILU)
    {
        if (!J->transpose)
        {
            for (int k = 0; k < nF; k++)
            {
                int sf = R->losort[k];
                wA[R->u[sf]] -= rD[R->u[sf]] * R->lower[sf] * wA[R->l[sf]];
            }
            for (int f = nF - 1; f >= 0; f--) wA[R->l[f]] -= rD[R->l[f]] * R->upper[f] * wA[R->u[f]];
        }
        else
        {
            /* DILUPreconditioner::preconditionT: roles of upper/lower swapped */
            for (int f = 0; f < nF; f++) wA[R->u[f]] -= rD[R->u[f]] * R->upper[f] * wA[R->l[f]];
            for (int k = nF - 1; k >= 0; k--)
            {
                int sf = R->losort[k];
                wA[R->l[sf]] -= rD[R->l[sf]] * R->lower[sf] * wA[R->u[sf]];
            }
        }
    }
}
static int precondition_impl(orc_sys* s, const double* rIn, double* wOut, int transpose)
{
    if (s->precond < 0) return -1;
    pre_job J = {s, rIn, wOut, transpose};
    par_for(s->nRows, row_precondition, &J);
    return 0;
}

int orc_precondition(orc_sys* s, const double* r, double* w) { return precondition_impl(s, r, w, 0); }
int orc_preconditionT(orc_sys* s, const double* r, double* w) { return precondition_impl(s, r, w, 1); }

/* ------------------------------------------------------------------ solve */

static int orc_stop(const orc_opts* o, orc_perf* p)
{
    /* lduMatrix::solver::stop + lduSolverPerformance::checkConvergence */
    if (p->nIterations < o->minIter) return 0;
    if (p->finalResidual < o->tolerance ||
        (o->relTol > ORC_SMALL_ && p->finalResidual <= o->relTol * p->initialResidual))
        p->converged = 1;
    else
        p->converged = 0;
    if (p->nIterations >= o->maxIter || p->converged) return 1;
    return 0;
}

static int check_singularity(orc_perf* p, double residual)
{
    p->singular = !(residual > ORC_VSMALL);
    return p->singular;
}

/* coupledIterativeSolver::normFactor / lduMatrix::solver::normFactor (foam-extend:
 * full Amul with the mean value, HJ 5/Nov/2007) */
static double norm_factor(orc_sys* s, const double* x, const double* b, const double* Ax, double* tmp,
                          double* xRefField)
{
    double xRef = gsum(s, x) / (double)s->total; /* gAverage */
    for (int i = 0; i < s->total; i++) xRefField[i] = xRef;
    orc_amul(s, xRefField, tmp);
    return gsum_normterms(s, Ax, b, tmp) + ORC_SMALL;
}

#define HIST(k, v)                                   \
    do                                               \
    {                                                \
        if (history && (k) < historyCap) history[(k)] = (v); \
    } while (0)

static int solve_pcg(orc_sys* s, const orc_opts* o, double* x, const double* b, orc_perf* perf,
                     double* history, int historyCap)
{
    const int n = s->total;
    double* pA = (double*)malloc(sizeof(double) * (size_t)n);
    double* wA = (double*)malloc(sizeof(double) * (size_t)n);
    double* rA = (double*)malloc(sizeof(double) * (size_t)n);
    double* xRefF = (double*)malloc(sizeof(double) * (size_t)n);
    orc_amul(s, x, wA);
    for (int i = 0; i < n; i++) rA[i] = b[i] - wA[i];
    double nf = norm_factor(s, x, b, wA, pA, xRefF);
    perf->normFactor = nf;
    perf->initialResidual = orc_gsummag(s, rA) / nf;
    perf->finalResidual = perf->initialResidual;
    HIST(0, perf->initialResidual);
    if (!orc_stop(o, perf))
    {
        double wArA = ORC_GREAT, wArAold = wArA;
        orc_precond_setup(s, o->precond);
        do
        {
            wArAold = wArA;
            orc_precondition(s, rA, wA);
            wArA = orc_gsumprod(s, wA, rA);
            if (perf->nIterations == 0)
                for (int i = 0; i < n; i++) pA[i] = wA[i];
            else
            {
                double beta = wArA / wArAold;
                for (int i = 0; i < n; i++) pA[i] = wA[i] + beta * pA[i];
            }
            orc_amul(s, pA, wA);
            double wApA = orc_gsumprod(s, wA, pA);
            if (check_singularity(perf, fabs(wApA) / nf)) break;
            double alpha = wArA / wApA;
            for (int i = 0; i < n; i++)
            {
                x[i] += alpha * pA[i];
                rA[i] -= alpha * wA[i];
            }
            perf->finalResidual = orc_gsummag(s, rA) / nf;
            perf->nIterations++;
            HIST(perf->nIterations, perf->finalResidual);
        } while (!orc_stop(o, perf));
    }
    free(pA);
    free(wA);
    free(rA);
    free(xRefF);
    return 0;
}

static int solve_bicgstab(orc_sys* s, const orc_opts* o, double* x, const double* b, orc_perf* perf,
                          double* history, int historyCap)
{
    /* bicgStabSolver::solve / coupledBicgStabSolver::solve */
    const int n = s->total;
    size_t nb = sizeof(double) * (size_t)(n ? n : 1);
    double* p = (double*)malloc(nb);
    double* r = (double*)malloc(nb);
    double* tmp = (double*)malloc(nb);
    double* xRefF = (double*)malloc(nb);
    /* normFactor(x, b): own Amul of x */
    orc_amul(s, x, p);
    double nf = norm_factor(s, x, b, p, tmp, xRefF);
    perf->normFactor = nf;
    orc_amul(s, x, p);
    for (int i = 0; i < n; i++) r[i] = b[i] - p[i];
    perf->initialResidual = orc_gsummag(s, r) / nf;
    perf->finalResidual = perf->initialResidual;
    HIST(0, perf->initialResidual);
    if (!orc_stop(o, perf))
    {
        double rho = ORC_GREAT, rhoOld = rho, alpha = 0, omega = ORC_GREAT, beta;
        double* ph = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
        double* v = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
        double* sv = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
        double* sh = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
        double* t = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
        double* rw = (double*)malloc(nb);
        for (int i = 0; i < n; i++) p[i] = 0.0;
        for (int i = 0; i < n; i++) rw[i] = r[i];
        orc_precond_setup(s, o->precond);
        do
        {
            rhoOld = rho;
            rho = orc_gsumprod(s, rw, r);
            beta = rho / rhoOld * (alpha / omega);
            if (rho == 0)
            {
                /* restart if breakdown occurs */
                for (int i = 0; i < n; i++) rw[i] = r[i];
                rho = orc_gsumprod(s, rw, r);
                alpha = 0;
                omega = 0;
                beta = 0;
            }
            vec_op(0, n, p, r, v, beta, omega);
            orc_precondition(s, p, ph);
            orc_amul(s, ph, v);
            alpha = rho / orc_gsumprod(s, rw, v);
            vec_op(1, n, sv, r, v, alpha, 0.0);
            orc_precondition(s, sv, sh);
            orc_amul(s, sh, t);
            omega = orc_gsumprod(s, t, sv) / orc_gsumprod(s, t, t);
            vec_op(2, n, x, ph, sh, alpha, omega);
            vec_op(3, n, r, sv, t, omega, 0.0);
            perf->finalResidual = orc_gsummag(s, r) / nf;
            perf->nIterations++;
            HIST(perf->nIterations, perf->finalResidual);
        } while (!orc_stop(o, perf));
        free(ph);
        free(v);
        free(sv);
        free(sh);
        free(t);
        free(rw);
    }
    free(p);
    free(r);
    free(tmp);
    free(xRefF);
    return 0;
}

static int solve_pbicg(orc_sys* s, const orc_opts* o, double* x, const double* b, orc_perf* perf,
                       double* history, int historyCap)
{
    /* PBiCG::solve (foam/matrices/lduMatrix/solvers/PBiCG/PBiCG.C) */
    const int n = s->total;
    size_t nb = sizeof(double) * (size_t)(n ? n : 1);
    double* pA = (double*)malloc(nb);
    double* pT = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
    double* wA = (double*)malloc(nb);
    double* wT = (double*)malloc(nb);
    double* rA = (double*)malloc(nb);
    double* rT = (double*)malloc(nb);
    double* xRefF = (double*)malloc(nb);
    orc_amul(s, x, wA);
    orc_tmul(s, x, wT);
    for (int i = 0; i < n; i++) rA[i] = b[i] - wA[i];
    for (int i = 0; i < n; i++) rT[i] = b[i] - wT[i];
    double nf = norm_factor(s, x, b, wA, pA, xRefF);
    perf->normFactor = nf;
    perf->initialResidual = orc_gsummag(s, rA) / nf;
    perf->finalResidual = perf->initialResidual;
    HIST(0, perf->initialResidual);
    if (!orc_stop(o, perf))
    {
        double wArT = ORC_GREAT, wArTold = wArT;
        orc_precond_setup(s, o->precond);
        do
        {
            wArTold = wArT;
            orc_precondition(s, rA, wA);
            orc_preconditionT(s, rT, wT);
            wArT = orc_gsumprod(s, wA, rT);
            if (perf->nIterations == 0)
            {
                for (int i = 0; i < n; i++)
                {
                    pA[i] = wA[i];
                    pT[i] = wT[i];
                }
            }
            else
            {
                double beta = wArT / wArTold;
                for (int i = 0; i < n; i++)
                {
                    pA[i] = wA[i] + beta * pA[i];
                    pT[i] = wT[i] + beta * pT[i];
                }
            }
            orc_amul(s, pA, wA);
            orc_tmul(s, pT, wT);
            double wApT = orc_gsumprod(s, wA, pT);
            if (check_singularity(perf, fabs(wApT) / nf)) break;
            double alpha = wArT / wApT;
            for (int i = 0; i < n; i++)
            {
                x[i] += alpha * pA[i];
                rA[i] -= alpha * wA[i];
                rT[i] -= alpha * wT[i];
            }
            perf->finalResidual = orc_gsummag(s, rA) / nf;
            perf->nIterations++;
            HIST(perf->nIterations, perf->finalResidual);
        } while (!orc_stop(o, perf));
    }
    free(pA);
    free(pT);
    free(wA);
    free(wT);
    free(rA);
    free(rT);
    free(xRefF);
    return 0;
}

int orc_solve(orc_sys* s, const orc_opts* o, double* x, const double* b, orc_perf* perf,
              double* history, int historyCap)
{
    memset(perf, 0, sizeof(*perf));
    for (int r = 0; r < s->nRows; r++)
        if (!s->rows[r].diag) return -2;
    switch (o->solver)
    {
        case ORC_SOLVER_PCG: return solve_pcg(s, o, x, b, perf, history, historyCap);
        case ORC_SOLVER_BICGSTAB: return solve_bicgstab(s, o, x, b, perf, history, historyCap);
        case ORC_SOLVER_PBICG: return solve_pbicg(s, o, x, b, perf, history, historyCap);
        default: return -1;
    }
}

/* ------------------------------------------- partitioned face transfer */

void orc_patch_face_to_global(int nRanks, const int* pieceOffsets, const int* faceToGlobalAddr,
                              const double* pField, int nComp, int nZoneFaces, double* gField)
{
    /* globalPolyPatchTemplates.C:162-179: every rank zero-fills a zone-sized
     * field, scatters its own patch values, then reduce(sum) over ranks. */
    double* piece = (double*)malloc(sizeof(double) * (size_t)(nZoneFaces * nComp + 1));
    for (int i = 0; i < nZoneFaces * nComp; i++) gField[i] = 0.0;
    for (int k = 0; k < nRanks; k++)
    {
        for (int i = 0; i < nZoneFaces * nComp; i++) piece[i] = 0.0;
        for (int i = pieceOffsets[k]; i < pieceOffsets[k + 1]; i++)
            for (int d = 0; d < nComp; d++) piece[faceToGlobalAddr[i] * nComp + d] = pField[i * nComp + d];
        if (k == 0)
            for (int i = 0; i < nZoneFaces * nComp; i++) gField[i] = piece[i];
        else
            for (int i = 0; i < nZoneFaces * nComp; i++) gField[i] += piece[i];
    }
    free(piece);
}

void orc_global_face_to_patch(int nLocal, const int* faceToGlobalAddr, const double* gField,
                              int nComp, double* pField)
{
    for (int i = 0; i < nLocal; i++)
        for (int d = 0; d < nComp; d++) pField[i * nComp + d] = gField[faceToGlobalAddr[i] * nComp + d];
}

/* directMapInterfaceToInterfaceMapping::calcZoneAToZoneBFaceMap, the N^2 search as written in the reference
 * (src/numerics/interfaceToInterfaceMappings/directMapInterfaceToInterfaceMapping/directMapInterfaceToInterfaceMapping.C:
 * 155-168; the same loop builds the B-to-A face map :276-289 and the two point maps :397-410, :518-531):
 *   map = -1;  forAll(to, i) forAll(from, j) if (mag(to[i] - from[j]) < tol) { map[i] = j; break; }
 * Returns the number of entries left at -1 (the reference is fatal when gMin(map) == -1, :170-181). */
int orc_direct_map_build(int nTo, const double* to, int nFrom, const double* from, double tol, int* map)
{
    int unmatched = 0;
    for (int i = 0; i < nTo; i++)
    {
        map[i] = -1;
        for (int j = 0; j < nFrom; j++)
        {
            const double dx = to[3 * i] - from[3 * j], dy = to[3 * i + 1] - from[3 * j + 1], dz = to[3 * i + 2] - from[3 * j + 2];
            if (sqrt(dx * dx + dy * dy + dz * dz) < tol)
            {
                map[i] = j;
                break;
            }
        }
        unmatched += map[i] < 0;
    }
    return unmatched;
}

void orc_direct_map(int nTo, const int* map, const double* from, int nComp, double* to)
{
    for (int i = 0; i < nTo; i++)
        for (int d = 0; d < nComp; d++) to[i * nComp + d] = from[map[i] * nComp + d];
}

/* ------------------------------------------------------------------------------------------------------------------
 * GaussSeidelSmoother / smoothSolver (`smoothSolver` with `smoother GaussSeidel` in the tutorials'
 * fvSolution, e.g. tutorials/fluidStructureInteraction/HronTurekFsi3/system/fluid/fvSolution).  foam-extend 4.1
 * foam/matrices/lduMatrix/smoothers/GaussSeidel/GaussSeidelSmoother.C and solvers/smoothSolver/smoothSolver.C, restated for
 * ONE matrix whose coupled-patch contributions are already inside bPrime (the smoother adds them to a copy of the source
 * before every sweep: bPrime = source; updateMatrixInterfaces(-bouCoeffs ...)).  Plain arrays, no orc_sys. */
void orc_gs_smooth(int n, int nf, const int* l, const int* u, const double* diag, const double* upper, const double* lower,
                   double* psi, const double* source, int nSweeps)
{
    if (!lower) lower = upper;
    int* ownStart = (int*)calloc((size_t)n + 1, sizeof(int));
    for (int f = 0; f < nf; f++) ownStart[l[f] + 1]++;
    for (int c = 0; c < n; c++) ownStart[c + 1] += ownStart[c];
    double* bPrime = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1));
    for (int sweep = 0; sweep < nSweeps; sweep++)
    {
        for (int c = 0; c < n; c++) bPrime[c] = source[c];
        for (int c = 0; c < n; c++)
        {
            const int fStart = ownStart[c], fEnd = ownStart[c + 1];
            double curPsi = bPrime[c];
            for (int f = fStart; f < fEnd; f++) curPsi -= upper[f] * psi[u[f]];
            curPsi /= diag[c];
            for (int f = fStart; f < fEnd; f++) bPrime[u[f]] -= lower[f] * curPsi;
            psi[c] = curPsi;
        }
    }
    free(bPrime);
    free(ownStart);
}

static void gs_amul(int n, int nf, const int* l, const int* u, const double* diag, const double* upper, const double* lower,
                    const double* x, double* y)
{
    for (int c = 0; c < n; c++) y[c] = diag[c] * x[c];
    for (int f = 0; f < nf; f++)
    {
        y[u[f]] += lower[f] * x[l[f]];
        y[l[f]] += upper[f] * x[u[f]];
    }
}

/* smoothSolver::solve with nSweeps > 0.  history[k] = residual after k rounds of nSweeps sweeps (entry 0 = initial). */
int orc_gs_solve(int n, int nf, const int* l, const int* u, const double* diag, const double* upper, const double* lower,
                 double* psi, const double* source, int nSweeps, double tolerance, double relTol, int minIter, int maxIter,
                 orc_perf* perf, double* history, int historyCap)
{
    if (nSweeps <= 0) return -1;
    if (!lower) lower = upper;
    memset(perf, 0, sizeof(*perf));
    double* Ax = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1));
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1));
    gs_amul(n, nf, l, u, diag, upper, lower, psi, Ax);
    double xRef = 0.0;
    for (int c = 0; c < n; c++) xRef += psi[c];
    xRef /= (double)(n ? n : 1);
    for (int c = 0; c < n; c++) tmp[c] = xRef;
    double* pA = (double*)malloc(sizeof(double) * (size_t)(n ? n : 1));
    gs_amul(n, nf, l, u, diag, upper, lower, tmp, pA);
    double nfac = 0.0;
    for (int c = 0; c < n; c++) nfac += fabs(Ax[c] - pA[c]) + fabs(source[c] - pA[c]);
    nfac += ORC_SMALL;
    perf->normFactor = nfac;
    double s0 = 0.0;
    for (int c = 0; c < n; c++) s0 += fabs(source[c] - Ax[c]);
    perf->initialResidual = perf->finalResidual = s0 / nfac;
    int k = 0;
    if (history && k < historyCap) history[k] = perf->initialResidual;
    orc_opts o;
    memset(&o, 0, sizeof(o));
    o.tolerance = tolerance;
    o.relTol = relTol;
    o.minIter = minIter;
    o.maxIter = maxIter;
    if (!orc_stop(&o, perf))
    {
        do
        {
            orc_gs_smooth(n, nf, l, u, diag, upper, lower, psi, source, nSweeps);
            /* lduMatrix::residual: rA = source - diag*psi; rA[u] -= lower*psi[l]; rA[l] -= upper*psi[u] */
            for (int c = 0; c < n; c++) tmp[c] = source[c] - diag[c] * psi[c];
            for (int f = 0; f < nf; f++)
            {
                tmp[u[f]] -= lower[f] * psi[l[f]];
                tmp[l[f]] -= upper[f] * psi[u[f]];
            }
            double s = 0.0;
            for (int c = 0; c < n; c++) s += fabs(tmp[c]);
            perf->finalResidual = s / nfac;
            perf->nIterations += nSweeps;
            k++;
            if (history && k < historyCap) history[k] = perf->finalResidual;
        } while (!orc_stop(&o, perf));
    }
    free(Ax);
    free(tmp);
    free(pA);
    return 0;
}
