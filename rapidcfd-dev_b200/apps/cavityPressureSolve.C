// cavityPressureSolve -- mini-application over the OpenFOAM-shaped facade (foam/b200Foam.H).
//
// Stands in for the pressure-corrector part of the reference's icoFoam time step
// (applications/solvers/incompressible/icoFoam/icoFoam.C:86-92: pEqn(fvm::laplacian(rAU,p)
// == fvc::div(phiHbyA)); pEqn.solve()) on the synthetic n^3 hex cavity: it assembles the
// pressure Laplacian with the device fill kernel (gaussLaplacianScheme.C:63-64), pins the
// reference cell (fvMatrix::setReference, fvMatrix.C:965-983), reads an fvSolution-style
// dictionary, selects the solver through lduMatrix::solver::New and prints the reference's
// solverPerformance line.
//
//   cavityPressureSolve <n> [fvSolution-file]      (default dictionary: PCG + DIC, 1e-6)
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../foam/b200Foam.H"

using namespace b200;

// A solver type added from OUTSIDE the facade, the way a user library registers one in the reference
// (PCG.C:36-37: lduMatrix::solver::addsymMatrixConstructorToTable<PCG>): `solver loggedPCG;` in the dictionary
// selects it through lduMatrix::solver::New.  It runs the library's PCG and reports under its own name.
class loggedPCG : public lduMatrix::solver
{
public:
    loggedPCG(const word &fieldName, const lduMatrix &m, const dictionary &d) : lduMatrix::solver(fieldName, m, d, "PCG") {}
    virtual solverPerformance solve(scalargpuField &psi, const scalargpuField &source, const int cmpt = 0) const
    {
        solverPerformance sp = lduMatrix::solver::solve(psi, source, cmpt);
        std::cout << "loggedPCG: " << sp.nIterations() << " iterations" << std::endl;
        return sp;
    }
};
static lduMatrix::solver::addsymMatrixConstructorToTable<loggedPCG> addloggedPCGSymMatrixConstructorToTable_("loggedPCG");

static const char *defaultDict =
    "solvers\n{\n    p\n    {\n        solver          PCG;\n        preconditioner  DIC;\n"
    "        tolerance       1e-06;\n        relTol          0;\n    }\n}\n";

int main(int argc, char **argv)
{
    try {
        const int n = argc > 1 ? atoi(argv[1]) : 32;
        std::string text = defaultDict;
        if (argc > 2) {
            std::ifstream f(argv[2]);
            if (!f) throw FatalError(std::string("cannot open ") + argv[2]);
            std::stringstream ss;
            ss << f.rdbuf();
            text = ss.str();
        }
        dictionary fvSolution(text);
        const dictionary &pDict = fvSolution.subDict("solvers").subDict("p");

        // ---- blockMesh stand-in: n^3 hex cavity in OpenFOAM ordering ----
        const double h = 1.0 / n;
        const int N = n * n * n;
        std::vector<label> lower, upper;
        std::vector<scalar> centres(3 * (size_t)N), weights;
        for (int k = 0; k < n; k++)
            for (int j = 0; j < n; j++)
                for (int i = 0; i < n; i++) {
                    int c = i + n * (j + n * k);
                    centres[3 * (size_t)c] = (i + 0.5) * h;
                    centres[3 * (size_t)c + 1] = (j + 0.5) * h;
                    centres[3 * (size_t)c + 2] = (k + 0.5) * h;
                    if (i < n - 1) {
                        lower.push_back(c);
                        upper.push_back(c + 1);
                        weights.push_back(h);
                    }
                    if (j < n - 1) {
                        lower.push_back(c);
                        upper.push_back(c + n);
                        weights.push_back(h * 1.01);
                    }
                    if (k < n - 1) {
                        lower.push_back(c);
                        upper.push_back(c + n * n);
                        weights.push_back(h * 1.02);
                    }
                }
        const int F = (int)lower.size();

        Context ctx(0);
        lduAddressing addr(ctx, N, lower, upper, {}, {}, {}, centres);
        lduMatrix pEqn(addr);

        // fvm::laplacian(rAU, p): upper = deltaCoeffs * (rAU*magSf), diag = -sum (device fill)
        scalargpuField deltaCoeffs(std::vector<scalar>(F, 1.0 / h));
        std::vector<scalar> gam(F);
        for (int f = 0; f < F; f++) gam[f] = (0.75 + 0.5 * ((f * 2654435761u) % 1000) / 1000.0) * h * h;
        scalargpuField gammaMagSf(gam);
        pEqn.upper().setSize(F);
        pEqn.diag().setSize(N);
        check(b200ldu_fv_laplacian_fill(addr.handle(), deltaCoeffs.data(), gammaMagSf.data(), pEqn.upper().data(),
                                        pEqn.diag().data()),
              "fvm::laplacian");
        { // setReference(pRefCell 0, pRefValue 0): diag[0] += diag[0]
            std::vector<scalar> d = pEqn.diag().toHost();
            d[0] += d[0];
            pEqn.diag() = d;
        }
        // source = fvc::div(phiHbyA) stand-in: zero-mean oscillating field
        std::vector<scalar> src(N);
        double mean = 0;
        for (int c = 0; c < N; c++) {
            src[c] = std::sin(12.9898 * c) * h * h * h;
            mean += src[c];
        }
        for (int c = 0; c < N; c++) src[c] -= mean / N;
        scalargpuField source(src), p((size_t)N, 0.0);

        std::unique_ptr<lduMatrix::solver> solverPtr = lduMatrix::solver::New("p", pEqn, pDict);
        if (pDict.lookup("solver") == "GAMG") solverPtr->setAgglomeration(weights);
        solverPerformance sp = solverPtr->solve(p, source);
        sp.print(std::cout);

        // independent check of the returned field: || source - A p ||_1 via lduMatrix::residual
        scalargpuField rA((size_t)N);
        pEqn.residual(rA, p, source);
        double r1 = 0, s1 = 0;
        std::vector<scalar> r = rA.toHost();
        for (int c = 0; c < N; c++) {
            r1 += std::fabs(r[c]);
            s1 += std::fabs(src[c]);
        }
        std::cout << "cells " << N << "  |b - A p|_1 / |b|_1 = " << r1 / s1 << "  kernel launches " << ctx.launches()
                  << std::endl;
        return sp.converged() || sp.nIterations() > 0 ? 0 : 1;
    } catch (const FatalError &e) {
        std::cerr << e.what() << std::endl;
        return 2;
    }
}
