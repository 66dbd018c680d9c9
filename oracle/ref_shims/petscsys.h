/* Single-rank stand-in for the slice of PETSc the reference's GMG path touches
 * (Vec, IS, VecScatter, AO, a minimal assembled Mat; PC/KSP only as link stubs).  TEST INFRASTRUCTURE.
 * Semantics follow SURVEY.md App. C: new Vecs are zero-initialised; block scatters are
 * FORWARD/INSERT y[j*bs+k] = x[idx[j]*bs+k] and REVERSE/ADD x[idx[j]*bs+k] += y[j*bs+k]. */
#ifndef ORACLE_SHIM_PETSCSYS_H
#define ORACLE_SHIM_PETSCSYS_H
#include <cstddef>
#include <cstdlib>
#include <map>
#include <vector>
#include <mpi.h>
using std::size_t;
typedef int    PetscInt;
typedef int    PetscErrorCode;
typedef double PetscScalar;
typedef double PetscReal;
typedef int    PetscMPIInt;
typedef bool   PetscBool;
#define PETSC_COMM_WORLD MPI_COMM_WORLD
#define PETSC_COMM_SELF MPI_COMM_SELF
#define PETSC_DETERMINE (-1)
#define PETSC_DECIDE (-1)
#define PETSC_TRUE true
#define PETSC_FALSE false
enum PetscCopyMode { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER };
enum InsertMode { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES };
enum ScatterMode { SCATTER_FORWARD, SCATTER_REVERSE };
struct _p_PetscObject { virtual ~_p_PetscObject() {} };
typedef _p_PetscObject *PetscObject;
struct _p_Vec : _p_PetscObject { std::vector<double> data; };
typedef _p_Vec *Vec;
struct _p_IS : _p_PetscObject { int bs; std::vector<int> idx; };
typedef _p_IS *IS;
struct _p_VecScatter : _p_PetscObject { int bs; std::vector<int> idx; };
typedef _p_VecScatter *VecScatter;
struct _p_AO : _p_PetscObject { std::map<int, int> map; };
typedef _p_AO *AO;
struct _p_Mat : _p_PetscObject { std::vector<std::map<int, double>> rows; /* assembled rows (MatSetValues), single rank */ };
typedef _p_Mat *Mat;
struct _p_PC : _p_PetscObject {};
typedef _p_PC *PC;
struct _p_KSP : _p_PetscObject {};
typedef _p_KSP *KSP;
typedef const char *MatType;
typedef const char *PCType;
#define MATMPIAIJ "mpiaij"
#define MATAIJ "aij"
#define PCSHELL "shell"
enum MatAssemblyType { MAT_FINAL_ASSEMBLY, MAT_FLUSH_ASSEMBLY };

PetscErrorCode PetscObjectDestroy(PetscObject *);
PetscErrorCode PetscInitialize(int *, char ***, const char *, const char *);
PetscErrorCode PetscFinalize();
/* Vec */
PetscErrorCode VecCreateMPI(MPI_Comm, PetscInt n, PetscInt N, Vec *);
PetscErrorCode VecCreateSeq(MPI_Comm, PetscInt n, Vec *);
PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecGetArray(Vec, double **);
PetscErrorCode VecGetArrayRead(Vec, const double **);
PetscErrorCode VecRestoreArray(Vec, double **);
PetscErrorCode VecRestoreArrayRead(Vec, const double **);
PetscErrorCode VecGetLocalSize(Vec, PetscInt *);
PetscErrorCode VecSet(Vec, double);
PetscErrorCode VecScale(Vec, double);
PetscErrorCode VecShift(Vec, double);
PetscErrorCode VecAXPY(Vec y, double a, Vec x);
PetscErrorCode VecAYPX(Vec y, double a, Vec x);
/* IS / VecScatter / AO */
PetscErrorCode ISCreateBlock(MPI_Comm, PetscInt bs, PetscInt n, const PetscInt idx[], PetscCopyMode, IS *);
PetscErrorCode VecScatterCreate(Vec x, IS ix, Vec y, IS iy, VecScatter *);
PetscErrorCode VecScatterBegin(VecScatter, Vec from, Vec to, InsertMode, ScatterMode);
PetscErrorCode VecScatterEnd(VecScatter, Vec from, Vec to, InsertMode, ScatterMode);
PetscErrorCode AOCreateMapping(MPI_Comm, PetscInt n, const PetscInt app[], const PetscInt petsc[], AO *);
PetscErrorCode AOApplicationToPetsc(AO, PetscInt n, PetscInt ia[]);
/* Mat: a row-map sparse matrix, enough for MatrixHelper::formCRSMatrix + MatMult (the assembled-operator cross-check of
 * ref_driver's `matapply`); the matrix-free GMG path never reaches it.  PC: link stubs only */
PetscErrorCode MatCreate(MPI_Comm, Mat *);
PetscErrorCode MatSetSizes(Mat, PetscInt, PetscInt, PetscInt, PetscInt);
PetscErrorCode MatSetType(Mat, MatType);
PetscErrorCode MatMPIAIJSetPreallocation(Mat, PetscInt, const PetscInt *, PetscInt, const PetscInt *);
PetscErrorCode MatSetValues(Mat, PetscInt, const PetscInt *, PetscInt, const PetscInt *, const PetscScalar *, InsertMode);
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType);
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType);
PetscErrorCode MatMult(Mat, Vec, Vec);
PetscErrorCode MatGetRow(Mat, PetscInt, PetscInt *, const PetscInt **, const PetscScalar **);
PetscErrorCode MatRestoreRow(Mat, PetscInt, PetscInt *, const PetscInt **, const PetscScalar **);
PetscErrorCode PCSetType(PC, PCType);
PetscErrorCode PCShellSetContext(PC, void *);
PetscErrorCode PCShellGetContext(PC, void **);
PetscErrorCode PCShellSetApply(PC, PetscErrorCode (*)(PC, Vec, Vec));
#endif
