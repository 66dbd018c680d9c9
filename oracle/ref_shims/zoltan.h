/* Single-rank stand-in for <zoltan.h>: partitioning yields "no imports, no exports".
 * TEST INFRASTRUCTURE (oracle/_ref build only). */
#ifndef ORACLE_SHIM_ZOLTAN_H
#define ORACLE_SHIM_ZOLTAN_H
#include <mpi.h>
typedef unsigned int  ZOLTAN_ID_TYPE;
typedef ZOLTAN_ID_TYPE *ZOLTAN_ID_PTR;
#define ZOLTAN_OK 0
#define ZOLTAN_COMPRESSED_VERTEX 2
#define ZOLTAN_COMPRESSED_EDGE 1
struct Zoltan_Struct { int unused; };
static inline Zoltan_Struct *Zoltan_Create(MPI_Comm) { return new Zoltan_Struct(); }
static inline void Zoltan_Destroy(Zoltan_Struct **zz) { delete *zz; *zz = nullptr; }
static inline int Zoltan_Set_Param(Zoltan_Struct *, const char *, const char *) { return ZOLTAN_OK; }
#define ORACLE_ZOLTAN_SETTER(name) \
	template <class F> static inline int name(Zoltan_Struct *, F, void *) { return ZOLTAN_OK; }
ORACLE_ZOLTAN_SETTER(Zoltan_Set_Num_Obj_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_Obj_List_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_HG_Size_CS_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_HG_CS_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_Obj_Size_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_Pack_Obj_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_Unpack_Obj_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_Num_Fixed_Obj_Fn)
ORACLE_ZOLTAN_SETTER(Zoltan_Set_Fixed_Obj_List_Fn)
static inline int Zoltan_LB_Partition(Zoltan_Struct *, int *changes, int *ngid, int *nlid, int *nimp,
                                      ZOLTAN_ID_PTR *ig, ZOLTAN_ID_PTR *il, int **ip, int **itp, int *nexp,
                                      ZOLTAN_ID_PTR *eg, ZOLTAN_ID_PTR *el, int **ep, int **etp)
{
	*changes = 0; *ngid = 1; *nlid = 0; *nimp = 0; *nexp = 0;
	*ig = *il = *eg = *el = nullptr; *ip = *itp = *ep = *etp = nullptr;
	return ZOLTAN_OK;
}
static inline int Zoltan_Migrate(Zoltan_Struct *, int, ZOLTAN_ID_PTR, ZOLTAN_ID_PTR, int *, int *, int,
                                 ZOLTAN_ID_PTR, ZOLTAN_ID_PTR, int *, int *) { return ZOLTAN_OK; }
#endif
