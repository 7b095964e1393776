/*
 * fv_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) of the T-equation assembly the reference performs every
 * time step before the coupled solve.  Only tests/ may load it; the product never does.
 *
 * PARITY UNPINNED: the operators below live in foam-extend 4.1 (finiteVolume/{EulerDdtScheme.C,
 * gaussConvectionScheme.C, gaussLaplacianScheme.C, fvMatrix.C}, lduMatrixOperations.C), which is not in
 * /root/reference; they are restated from the published algorithm, one fvMatrix per operator and the same
 * sequential face loops, and anchored on the reference's call sites:
 *   src/regions/conductTemperature/conductTemperature.C:135-142
 *       TEqn = ( fvm::ddt(rho_*cv_, T()) == fvm::laplacian(kappa_(), T(), "laplacian(k,T)") );
 *   src/regions/transportTemperature/transportTemperature.C:129-140
 *       TEqn = ( rho_*cp_*( fvm::ddt(T()) + fvm::div(phi_(), T()) ) == fvm::laplacian(kappa_(), T()) );
 * Schemes as in tutorials/conjugateHeatTransfer/flowOverHeatedPlate/system/{fluid,solid}/fvSchemes:
 * Euler, Gauss upwind, Gauss linear uncorrected.
 *
 * Build: gcc -O3 -ffp-contract=off (oracle/Makefile) -- every product rounded, as the reference build.
 */
#include <stdlib.h>
#include <string.h>

typedef struct
{
    int n, nf;
    double *diag, *upper, *lower, *source;
} fvm_t;

static void fvm_init(fvm_t* m, int n, int nf)
{
    m->n = n;
    m->nf = nf;
    m->diag = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    m->source = (double*)calloc(n > 0 ? n : 1, sizeof(double));
    m->upper = (double*)calloc(nf > 0 ? nf : 1, sizeof(double));
    m->lower = (double*)calloc(nf > 0 ? nf : 1, sizeof(double));
}

static void fvm_free(fvm_t* m)
{
    free(m->diag);
    free(m->source);
    free(m->upper);
    free(m->lower);
}

/* lduMatrix::negSumDiag */
static void neg_sum_diag(fvm_t* m, const int* l, const int* u)
{
    for (int f = 0; f < m->nf; f++)
    {
        m->diag[l[f]] -= m->lower[f];
        m->diag[u[f]] -= m->upper[f];
    }
}

/* fvMatrix::operator*=(dimensioned<scalar>) */
static void fvm_scale(fvm_t* m, double s)
{
    for (int c = 0; c < m->n; c++)
    {
        m->diag[c] *= s;
        m->source[c] *= s;
    }
    for (int f = 0; f < m->nf; f++)
    {
        m->upper[f] *= s;
        m->lower[f] *= s;
    }
}

/* a += sign*b */
static void fvm_add(fvm_t* a, const fvm_t* b, int subtract)
{
    for (int c = 0; c < a->n; c++)
    {
        if (subtract)
        {
            a->diag[c] -= b->diag[c];
            a->source[c] -= b->source[c];
        }
        else
        {
            a->diag[c] += b->diag[c];
            a->source[c] += b->source[c];
        }
    }
    for (int f = 0; f < a->nf; f++)
    {
        if (subtract)
        {
            a->upper[f] -= b->upper[f];
            a->lower[f] -= b->lower[f];
        }
        else
        {
            a->upper[f] += b->upper[f];
            a->lower[f] += b->lower[f];
        }
    }
}

/*
 * form 0: ddt(rhoC,T) == laplacian(kappa,T);  form 1: rhoC*(ddt(T) + div(phi,T)) == laplacian(kappa,T).
 * kappaFace / phi may be NULL (uniform kappa / no flux).  Boundary faces (bCells, bInt, bSrc) in patch order are
 * added last, as fvMatrix::solve does (addBoundaryDiag, addBoundarySource).  Outputs: diag, upper, lower, source.
 */
int fvo_assemble_T(int form, int n, int nf, const int* l, const int* u, double rhoC, double rDeltaT, double kappa,
                   const double* kappaFace, const double* V, const double* magSf, const double* deltaCoeffs,
                   const double* phi, const double* Told, int nB, const int* bCells, const double* bInt,
                   const double* bSrc, double* diag, double* upper, double* lower, double* source)
{
    fvm_t ddt, lap;
    fvm_init(&ddt, n, nf);
    fvm_init(&lap, n, nf);

    /* gaussLaplacianScheme::fvmLaplacianUncorrected: gammaMagSf = gamma_f*magSf */
    for (int f = 0; f < nf; f++)
    {
        const double gammaMagSf = (kappaFace ? kappaFace[f] : kappa) * magSf[f];
        lap.upper[f] = deltaCoeffs[f] * gammaMagSf;
        lap.lower[f] = lap.upper[f]; /* symmetric */
    }
    neg_sum_diag(&lap, l, u);

    if (form == 0)
    {
        /* EulerDdtScheme::fvmDdt(const dimensionedScalar& rho, vf) */
        for (int c = 0; c < n; c++)
        {
            ddt.diag[c] = rDeltaT * rhoC * V[c];
            ddt.source[c] = rDeltaT * rhoC * Told[c] * V[c];
        }
    }
    else
    {
        /* EulerDdtScheme::fvmDdt(vf) */
        for (int c = 0; c < n; c++)
        {
            ddt.diag[c] = rDeltaT * V[c];
            ddt.source[c] = rDeltaT * Told[c] * V[c];
        }
        /* gaussConvectionScheme::fvmDiv with upwind weights (pos(faceFlux)) */
        fvm_t div;
        fvm_init(&div, n, nf);
        if (phi)
        {
            for (int f = 0; f < nf; f++)
            {
                const double w = phi[f] >= 0.0 ? 1.0 : 0.0;
                div.lower[f] = -w * phi[f];
                div.upper[f] = div.lower[f] + phi[f];
            }
            neg_sum_diag(&div, l, u);
        }
        fvm_add(&ddt, &div, 0);  /* ddt + div */
        fvm_scale(&ddt, rhoC);   /* rho*cp*( ... ) */
        fvm_free(&div);
    }
    fvm_add(&ddt, &lap, 1); /* A == B  ->  A - B */

    /* fvMatrix::solve: addBoundaryDiag / addBoundarySource, patch by patch */
    for (int k = 0; k < nB; k++)
    {
        ddt.diag[bCells[k]] += bInt[k];
        ddt.source[bCells[k]] += bSrc[k];
    }
    memcpy(diag, ddt.diag, sizeof(double) * (size_t)n);
    memcpy(source, ddt.source, sizeof(double) * (size_t)n);
    memcpy(upper, ddt.upper, sizeof(double) * (size_t)nf);
    memcpy(lower, ddt.lower, sizeof(double) * (size_t)nf);
    fvm_free(&ddt);
    fvm_free(&lap);
    return 0;
}
