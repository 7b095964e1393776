/* blk_oracle.c -- CPU restatement of the block-coupled (vector4) solve path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this; the product never does.
 *
 * PARITY UNPINNED: the arithmetic lives in foam-extend 4.1, which is not in
 * /root/reference; the reference only reaches it through
 *     fvBlockMatrix<Type>::solve            /root/reference/filesToReplace/fvBlockMatrix.C:1360-1388
 *         -> BlockLduSolver<Type>::New(psi.name(), *this, dict)->solve(psi.internalField(), source())
 * with Type = vector4 (src/regions/pUCoupledIcoFluid/pUCoupledIcoFluid.C:584-621).
 * Restated from foam-extend 4.1:
 *   [FE] foam/matrices/blockLduMatrix/BlockLduMatrix/BlockLduMatrixATmul.C      Amul / AmulCore
 *   [FE] foam/matrices/blockLduMatrix/BlockLduMatrix/BlockCoeff*.H (multiply)   scalar / linear / square products
 *   [FE] foam/matrices/blockLduMatrix/BlockLduPrecons/BlockCholeskyPrecon       calcPreconDiag, ILUmultiply
 *   [FE] foam/matrices/blockLduMatrix/BlockLduPrecons/BlockDiagonalPrecon, BlockNoPrecon
 *   [FE] foam/matrices/blockLduMatrix/BlockLduSolvers/BlockBiCGStab, BlockCG, BlockIterativeSolver (normFactor, stop)
 * Details that could not be confirmed without that source are isolated and marked UNCONFIRMED:
 *   - inv(TensorN): restated as Gauss-Jordan with partial pivoting (blk_inv4);
 *   - mixed coefficient kinds in BlockCholesky are promoted to the widest kind before the triple product.
 *
 * Layout: x, b [N][4]; coefficients per entry SCALAR (1 double), LINEAR (4) or SQUARE (16, row-major (i,j)), chosen
 * per array exactly as CoeffField<vector4> does (fvBlockMatrix.C:84-126, 184-221).  lower == NULL: symmetric matrix,
 * the lower triangle is the TRANSPOSED upper coefficient (BlockLduMatrixATmul.C, "Use transpose upper coefficient").
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BLK_GREAT 1.0e+20
#define BLK_SMALL 1.0e-20
#define BLK_SMALL_ 1.0e-15

typedef struct blk_opts
{
    int solver;  /* 0 CG, 1 BiCGStab */
    int precond; /* 0 none, 1 diagonal, 4 Cholesky */
    double tolerance, relTol;
    int minIter, maxIter;
} blk_opts;

typedef struct blk_perf
{
    double initialResidual[4], finalResidual[4];
    int nIterations, converged, singular;
    double normFactor;
} blk_perf;

typedef struct blk_sys
{
    int n, nf;
    int *l, *u, *losort, *ownerStart, *losortStart;
    int dK, uK, lK; /* kinds; lK == 0: symmetric */
    double *diag, *upper, *lower;
    int precond, pK;
    double* pD;
    int redMode; /* 0: sequential sums (the reference); 1: pairwise sums - measures the reference's own
                    sensitivity to its summation order (tests only) */
} blk_sys;

blk_sys* blk_create(int nCells, int nFaces, const int* l, const int* u)
{
    blk_sys* s = (blk_sys*)calloc(1, sizeof(blk_sys));
    s->n = nCells;
    s->nf = nFaces;
    s->l = (int*)malloc(sizeof(int) * (size_t)(nFaces + 1));
    s->u = (int*)malloc(sizeof(int) * (size_t)(nFaces + 1));
    memcpy(s->l, l, sizeof(int) * (size_t)nFaces);
    memcpy(s->u, u, sizeof(int) * (size_t)nFaces);
    /* lduAddressing::calcLosort: faces ordered by upper address, stable in the face index */
    s->losort = (int*)malloc(sizeof(int) * (size_t)(nFaces + 1));
    s->losortStart = (int*)calloc((size_t)nCells + 2, sizeof(int));
    s->ownerStart = (int*)calloc((size_t)nCells + 2, sizeof(int));
    for (int f = 0; f < nFaces; f++)
    {
        s->losortStart[u[f] + 1]++;
        s->ownerStart[l[f] + 1]++;
    }
    for (int c = 0; c < nCells; c++)
    {
        s->losortStart[c + 1] += s->losortStart[c];
        s->ownerStart[c + 1] += s->ownerStart[c];
    }
    int* fill = (int*)calloc((size_t)nCells + 1, sizeof(int));
    for (int f = 0; f < nFaces; f++) s->losort[s->losortStart[u[f]] + fill[u[f]]++] = f;
    free(fill);
    return s;
}

void blk_destroy(blk_sys* s)
{
    if (!s) return;
    free(s->l);
    free(s->u);
    free(s->losort);
    free(s->ownerStart);
    free(s->losortStart);
    free(s->diag);
    free(s->upper);
    free(s->lower);
    free(s->pD);
    free(s);
}

static double* dup(const double* p, size_t n)
{
    double* q = (double*)malloc(sizeof(double) * (n ? n : 1));
    if (n) memcpy(q, p, sizeof(double) * n);
    return q;
}

int blk_set_coeffs(blk_sys* s, int dK, const double* diag, int uK, const double* upper, int lK, const double* lower)
{
    if ((dK != 1 && dK != 4 && dK != 16) || (uK != 1 && uK != 4 && uK != 16)) return -1;
    if (lower && lK != uK) return -1; /* "Assuming lower and upper triangle have the same active type" */
    free(s->diag);
    free(s->upper);
    free(s->lower);
    s->dK = dK;
    s->uK = uK;
    s->lK = lower ? lK : 0;
    s->diag = dup(diag, (size_t)s->n * dK);
    s->upper = dup(upper, (size_t)s->nf * uK);
    s->lower = lower ? dup(lower, (size_t)s->nf * lK) : NULL;
    s->precond = -1;
    return 0;
}

/* BlockCoeff<Type>::multiply::operator()(coeff, x) for the three active types; a SQUARE coefficient is
 * (a & x)_i = sum_j a(i,j) x(j), summed left to right.  tr: use the transposed square coefficient. */
static void blk_mult(int kind, const double* a, int tr, const double* x, double* y)
{
    if (kind == 1)
        for (int i = 0; i < 4; i++) y[i] = a[0] * x[i];
    else if (kind == 4)
        for (int i = 0; i < 4; i++) y[i] = a[i] * x[i];
    else
        for (int i = 0; i < 4; i++)
        {
            double sum = (tr ? a[i] : a[4 * i]) * x[0];
            for (int j = 1; j < 4; j++) sum += (tr ? a[4 * j + i] : a[4 * i + j]) * x[j];
            y[i] = sum;
        }
}

/* BlockLduMatrix<Type>::Amul -> AmulCore (no coupled interfaces: the block systems of the reference are solved per
 * region through the partitioned path, multiRegionSystem.C:293) */
int blk_amul(blk_sys* s, const double* x, double* y)
{
    double t[4];
    for (int c = 0; c < s->n; c++) blk_mult(s->dK, s->diag + (size_t)c * s->dK, 0, x + 4 * (size_t)c, y + 4 * (size_t)c);
    /* lower multiplication */
    for (int f = 0; f < s->nf; f++)
    {
        if (s->lK)
            blk_mult(s->lK, s->lower + (size_t)f * s->lK, 0, x + 4 * (size_t)s->l[f], t);
        else
            blk_mult(s->uK, s->upper + (size_t)f * s->uK, 1, x + 4 * (size_t)s->l[f], t);
        for (int i = 0; i < 4; i++) y[4 * (size_t)s->u[f] + i] += t[i];
    }
    /* upper multiplication */
    for (int f = 0; f < s->nf; f++)
    {
        blk_mult(s->uK, s->upper + (size_t)f * s->uK, 0, x + 4 * (size_t)s->u[f], t);
        for (int i = 0; i < 4; i++) y[4 * (size_t)s->l[f] + i] += t[i];
    }
    return 0;
}

/* UNCONFIRMED stand-in for inv(TensorN<4>): Gauss-Jordan on [m | I], partial pivoting (first row of largest
 * magnitude), pivot row scaled by the reciprocal of the pivot.  Mirrored operation by operation on the device. */
void blk_inv4(const double* a, double* out)
{
    double m[4][4], r[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
        {
            m[i][j] = a[4 * i + j];
            r[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int k = 0; k < 4; k++)
    {
        int p = k;
        double best = fabs(m[k][k]);
        for (int q = k + 1; q < 4; q++)
            if (fabs(m[q][k]) > best)
            {
                best = fabs(m[q][k]);
                p = q;
            }
        if (p != k)
            for (int j = 0; j < 4; j++)
            {
                double tm = m[k][j];
                m[k][j] = m[p][j];
                m[p][j] = tm;
                tm = r[k][j];
                r[k][j] = r[p][j];
                r[p][j] = tm;
            }
        const double piv = 1.0 / m[k][k];
        for (int j = 0; j < 4; j++)
        {
            m[k][j] *= piv;
            r[k][j] *= piv;
        }
        for (int q = 0; q < 4; q++)
        {
            if (q == k) continue;
            const double f = m[q][k];
            for (int j = 0; j < 4; j++)
            {
                m[q][j] -= f * m[k][j];
                r[q][j] -= f * r[k][j];
            }
        }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[4 * i + j] = r[i][j];
}

/* expand a coefficient of kind k to SQUARE (scalar -> s I, linear -> diagonal); tr: transposed */
static void blk_expand16(int k, const double* a, int tr, double* o)
{
    for (int i = 0; i < 16; i++) o[i] = 0.0;
    if (k == 1)
        for (int i = 0; i < 4; i++) o[5 * i] = a[0];
    else if (k == 4)
        for (int i = 0; i < 4; i++) o[5 * i] = a[i];
    else
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) o[4 * i + j] = tr ? a[4 * j + i] : a[4 * i + j];
}

static void blk_matmul(const double* a, const double* b, double* o)
{ /* (a & b)(i,j) = sum_k a(i,k) b(k,j), left to right */
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
        {
            double sum = a[4 * i] * b[j];
            for (int k = 1; k < 4; k++) sum += a[4 * i + k] * b[4 * k + j];
            o[4 * i + j] = sum;
        }
}

/* BlockCholeskyPrecon<Type>::calcPreconDiag: preconDiag = diag;
 *   face order: preconDiag[u] -= tripleProduct(lower[f] (sym: upper[f].T()), preconDiag[l], upper[f]);  then inverted.
 * tripleProduct: scalar a*c/b, linear cmptDivide(cmptMultiply(a, c), b), square (a & inv(b)) & c.
 * BlockDiagonalPrecon: inverse of the diagonal only. */
int blk_precond_setup(blk_sys* s, int precond)
{
    if (!s->diag) return -1;
    if (s->precond == precond) return 0;
    free(s->pD);
    s->pD = NULL;
    s->precond = precond;
    if (precond == 0) return 0;
    const int chol = precond == 4;
    int pK = s->dK;
    if (chol && s->uK > pK) pK = s->uK;
    s->pK = pK;
    double* pD = (double*)malloc(sizeof(double) * (size_t)(s->n ? s->n : 1) * pK);
    s->pD = pD;
    for (int c = 0; c < s->n; c++)
    {
        const double* d = s->diag + (size_t)c * s->dK;
        if (pK == 16)
            blk_expand16(s->dK, d, 0, pD + 16 * (size_t)c);
        else if (pK == 4)
            for (int i = 0; i < 4; i++) pD[4 * (size_t)c + i] = s->dK == 1 ? d[0] : d[i];
        else
            pD[c] = d[0];
    }
    if (chol)
        for (int f = 0; f < s->nf; f++)
        {
            const double* up = s->upper + (size_t)f * s->uK;
            const double* lo = s->lK ? s->lower + (size_t)f * s->lK : up;
            const int l = s->l[f], u = s->u[f];
            if (pK == 1)
                pD[u] -= lo[0] * up[0] / pD[l];
            else if (pK == 4)
                for (int i = 0; i < 4; i++)
                {
                    const double a = s->uK == 1 ? lo[0] : lo[i], c = s->uK == 1 ? up[0] : up[i];
                    pD[4 * (size_t)u + i] -= (a * c) / pD[4 * (size_t)l + i];
                }
            else
            {
                double A[16], C[16], Bi[16], AB[16], T[16];
                blk_expand16(s->uK, lo, s->lK ? 0 : 1, A);
                blk_expand16(s->uK, up, 0, C);
                blk_inv4(pD + 16 * (size_t)l, Bi);
                blk_matmul(A, Bi, AB);
                blk_matmul(AB, C, T);
                for (int i = 0; i < 16; i++) pD[16 * (size_t)u + i] -= T[i];
            }
        }
    for (int c = 0; c < s->n; c++)
    {
        if (pK == 16)
        {
            double t[16];
            blk_inv4(pD + 16 * (size_t)c, t);
            memcpy(pD + 16 * (size_t)c, t, sizeof(t));
        }
        else
            for (int i = 0; i < pK; i++) pD[(size_t)pK * c + i] = 1.0 / pD[(size_t)pK * c + i];
    }
    return 0;
}

int blk_get_precon_diag(blk_sys* s, double* out, int* kind)
{
    if (!s->pD) return -1;
    memcpy(out, s->pD, sizeof(double) * (size_t)s->n * s->pK);
    *kind = s->pK;
    return 0;
}

/* BlockCholeskyPrecon<Type>::precondition -> ILUmultiply:
 *   x[i] = mult(dDiag[i], b[i]);
 *   forward, losort order: x[u] -= mult(dDiag[u], mult(lower[f] (sym: upper[f].T()), x[l]));
 *   backward, reverse face order: x[l] -= mult(dDiag[l], mult(upper[f], x[u])). */
int blk_precondition(blk_sys* s, const double* b, double* x)
{
    const int n = s->n;
    if (s->precond < 0) return -1;
    if (s->precond == 0)
    {
        memcpy(x, b, sizeof(double) * 4 * (size_t)n);
        return 0;
    }
    const int pK = s->pK;
    for (int c = 0; c < n; c++) blk_mult(pK, s->pD + (size_t)pK * c, 0, b + 4 * (size_t)c, x + 4 * (size_t)c);
    if (s->precond != 4) return 0;
    double t[4], w[4];
    for (int k = 0; k < s->nf; k++)
    {
        const int f = s->losort[k], l = s->l[f], u = s->u[f];
        if (s->lK)
            blk_mult(s->lK, s->lower + (size_t)f * s->lK, 0, x + 4 * (size_t)l, t);
        else
            blk_mult(s->uK, s->upper + (size_t)f * s->uK, 1, x + 4 * (size_t)l, t);
        blk_mult(pK, s->pD + (size_t)pK * u, 0, t, w);
        for (int i = 0; i < 4; i++) x[4 * (size_t)u + i] -= w[i];
    }
    for (int f = s->nf - 1; f >= 0; f--)
    {
        const int l = s->l[f], u = s->u[f];
        blk_mult(s->uK, s->upper + (size_t)f * s->uK, 0, x + 4 * (size_t)u, t);
        blk_mult(pK, s->pD + (size_t)pK * l, 0, t, w);
        for (int i = 0; i < 4; i++) x[4 * (size_t)l + i] -= w[i];
    }
    return 0;
}

void blk_set_reduction_mode(blk_sys* s, int mode) { s->redMode = mode; }

static double pairwise(const double* v, int n)
{
    if (n <= 8)
    {
        double t = 0.0;
        for (int i = 0; i < n; i++) t += v[i];
        return t;
    }
    return pairwise(v, n / 2) + pairwise(v + n / 2, n - n / 2);
}

/* gSumProd of two Field<vector4>: sum over cells of (a & b), each inner product summed left to right */
double blk_gsumprod(const blk_sys* s, const double* a, const double* b)
{
    double sum = 0.0;
    double* terms = s->redMode ? (double*)malloc(sizeof(double) * (size_t)(s->n ? s->n : 1)) : NULL;
    for (int c = 0; c < s->n; c++)
    {
        const double *p = a + 4 * (size_t)c, *q = b + 4 * (size_t)c;
        double d = p[0] * q[0];
        for (int i = 1; i < 4; i++) d += p[i] * q[i];
        if (terms)
            terms[c] = d;
        else
            sum += d;
    }
    if (terms)
    {
        sum = pairwise(terms, s->n);
        free(terms);
    }
    return sum;
}

/* gSum(cmptMag(r)) */
void blk_gsumcmptmag(const blk_sys* s, const double* r, double* out4)
{
    for (int i = 0; i < 4; i++) out4[i] = 0.0;
    for (int c = 0; c < s->n; c++)
        for (int i = 0; i < 4; i++) out4[i] += fabs(r[4 * (size_t)c + i]);
}

static double mag4(const double* v)
{
    double q = v[0] * v[0];
    for (int i = 1; i < 4; i++) q += v[i] * v[i];
    return sqrt(q);
}

/* BlockIterativeSolver<Type>::normFactor: xRef = gAverage(x); wA = A x; pA = A xRef;
 * normFactor = gSum(mag(wA - pA) + mag(b - pA)) + small_ (a scalar; the residual itself is Type-valued) */
double blk_norm_factor(blk_sys* s, const double* x, const double* b)
{
    const int n = s->n;
    double* wA = (double*)malloc(sizeof(double) * 4 * (size_t)(n ? n : 1));
    double* pA = (double*)malloc(sizeof(double) * 4 * (size_t)(n ? n : 1));
    double* xr = (double*)malloc(sizeof(double) * 4 * (size_t)(n ? n : 1));
    double xRef[4] = {0, 0, 0, 0};
    for (int c = 0; c < n; c++)
        for (int i = 0; i < 4; i++) xRef[i] += x[4 * (size_t)c + i];
    for (int i = 0; i < 4; i++) xRef[i] /= (double)(n ? n : 1);
    for (int c = 0; c < n; c++)
        for (int i = 0; i < 4; i++) xr[4 * (size_t)c + i] = xRef[i];
    blk_amul(s, x, wA);
    blk_amul(s, xr, pA);
    double sum = 0.0;
    for (int c = 0; c < n; c++)
    {
        double d1[4], d2[4];
        for (int i = 0; i < 4; i++)
        {
            d1[i] = wA[4 * (size_t)c + i] - pA[4 * (size_t)c + i];
            d2[i] = b[4 * (size_t)c + i] - pA[4 * (size_t)c + i];
        }
        sum += mag4(d1) + mag4(d2);
    }
    free(wA);
    free(pA);
    free(xr);
    return sum + BLK_SMALL;
}

static double cmax4(const double* v)
{
    double m = v[0];
    for (int i = 1; i < 4; i++)
        if (v[i] > m) m = v[i];
    return m;
}

/* BlockLduSolver::stop + BlockSolverPerformance<Type>::checkConvergence (on cmptMax of the residuals) */
static int blk_stop(const blk_opts* o, blk_perf* p)
{
    if (p->nIterations < o->minIter) return 0;
    const double fin = cmax4(p->finalResidual), ini = cmax4(p->initialResidual);
    p->converged = (fin < o->tolerance || (o->relTol > BLK_SMALL_ && fin <= o->relTol * ini)) ? 1 : 0;
    return (p->nIterations >= o->maxIter || p->converged) ? 1 : 0;
}

#define BHIST(it, res)                                                         \
    do                                                                         \
    {                                                                          \
        if (history && (it) < cap)                                             \
            for (int i_ = 0; i_ < 4; i_++) history[4 * (it) + i_] = (res)[i_]; \
    } while (0)

static void set_residual(blk_sys* s, const double* r, double nf, double* out4)
{
    blk_gsumcmptmag(s, r, out4);
    for (int i = 0; i < 4; i++) out4[i] /= nf;
}

/* BlockBiCGStabSolver<Type>::solve */
static int blk_solve_bicgstab(blk_sys* s, const blk_opts* o, double* x, const double* b, blk_perf* perf, double* history, int cap)
{
    const size_t n4 = 4 * (size_t)s->n, nb = sizeof(double) * (n4 ? n4 : 1);
    const double nf = blk_norm_factor(s, x, b);
    perf->normFactor = nf;
    double* p = (double*)malloc(nb);
    double* r = (double*)malloc(nb);
    blk_amul(s, x, p);
    for (size_t i = 0; i < n4; i++) r[i] = b[i] - p[i];
    set_residual(s, r, nf, perf->initialResidual);
    memcpy(perf->finalResidual, perf->initialResidual, sizeof(double) * 4);
    BHIST(0, perf->initialResidual);
    if (!blk_stop(o, perf))
    {
        double rho = BLK_GREAT, rhoOld = rho, alpha = 0, omega = BLK_GREAT, beta;
        double* ph = (double*)calloc(n4 ? n4 : 1, sizeof(double));
        double* v = (double*)calloc(n4 ? n4 : 1, sizeof(double));
        double* sv = (double*)calloc(n4 ? n4 : 1, sizeof(double));
        double* sh = (double*)calloc(n4 ? n4 : 1, sizeof(double));
        double* t = (double*)calloc(n4 ? n4 : 1, sizeof(double));
        double* rw = (double*)malloc(nb);
        memset(p, 0, nb);
        memcpy(rw, r, nb);
        blk_precond_setup(s, o->precond);
        do
        {
            rhoOld = rho;
            rho = blk_gsumprod(s, rw, r);
            beta = rho / rhoOld * (alpha / omega);
            if (rho == 0)
            { /* restart if breakdown occurs */
                memcpy(rw, r, nb);
                rho = blk_gsumprod(s, rw, r);
                alpha = 0;
                omega = 0;
                beta = 0;
            }
            for (size_t i = 0; i < n4; i++) p[i] = r[i] + beta * p[i] - beta * omega * v[i];
            blk_precondition(s, p, ph);
            blk_amul(s, ph, v);
            alpha = rho / blk_gsumprod(s, rw, v);
            for (size_t i = 0; i < n4; i++) sv[i] = r[i] - alpha * v[i];
            blk_precondition(s, sv, sh);
            blk_amul(s, sh, t);
            omega = blk_gsumprod(s, t, sv) / blk_gsumprod(s, t, t);
            for (size_t i = 0; i < n4; i++)
            {
                x[i] = x[i] + alpha * ph[i] + omega * sh[i];
                r[i] = sv[i] - omega * t[i];
            }
            set_residual(s, r, nf, perf->finalResidual);
            perf->nIterations++;
            BHIST(perf->nIterations, perf->finalResidual);
        } while (!blk_stop(o, perf));
        free(ph);
        free(v);
        free(sv);
        free(sh);
        free(t);
        free(rw);
    }
    free(p);
    free(r);
    return 0;
}

/* BlockCGSolver<Type>::solve (symmetric block systems) */
static int blk_solve_cg(blk_sys* s, const blk_opts* o, double* x, const double* b, blk_perf* perf, double* history, int cap)
{
    const size_t n4 = 4 * (size_t)s->n, nb = sizeof(double) * (n4 ? n4 : 1);
    const double nf = blk_norm_factor(s, x, b);
    perf->normFactor = nf;
    double* wA = (double*)malloc(nb);
    double* rA = (double*)malloc(nb);
    blk_amul(s, x, wA);
    for (size_t i = 0; i < n4; i++) rA[i] = b[i] - wA[i];
    set_residual(s, rA, nf, perf->initialResidual);
    memcpy(perf->finalResidual, perf->initialResidual, sizeof(double) * 4);
    BHIST(0, perf->initialResidual);
    if (!blk_stop(o, perf))
    {
        double rho = BLK_GREAT, rhoOld = rho;
        double* pA = (double*)calloc(n4 ? n4 : 1, sizeof(double));
        blk_precond_setup(s, o->precond);
        do
        {
            rhoOld = rho;
            blk_precondition(s, rA, wA);
            rho = blk_gsumprod(s, wA, rA);
            const double beta = rho / rhoOld;
            for (size_t i = 0; i < n4; i++) pA[i] = wA[i] + beta * pA[i];
            blk_amul(s, pA, wA);
            const double wApA = blk_gsumprod(s, wA, pA);
            /* checkSingularity(mag(wApA)/norm) */
            if (!(fabs(wApA) / nf > 1.0e-300))
            { /* VSMALL */
                perf->singular = 1;
                break;
            }
            const double alpha = rho / wApA;
            for (size_t i = 0; i < n4; i++)
            {
                x[i] += alpha * pA[i];
                rA[i] -= alpha * wA[i];
            }
            set_residual(s, rA, nf, perf->finalResidual);
            perf->nIterations++;
            BHIST(perf->nIterations, perf->finalResidual);
        } while (!blk_stop(o, perf));
        free(pA);
    }
    free(wA);
    free(rA);
    return 0;
}

int blk_solve(blk_sys* s, const blk_opts* o, double* x, const double* b, blk_perf* perf, double* history, int cap)
{
    if (!s->diag) return -1;
    memset(perf, 0, sizeof(*perf));
    if (history)
        for (int i = 0; i < 4 * cap; i++) history[i] = NAN;
    if (o->solver == 0) return blk_solve_cg(s, o, x, b, perf, history, cap);
    if (o->solver == 1) return blk_solve_bicgstab(s, o, x, b, perf, history, cap);
    return -1;
}
