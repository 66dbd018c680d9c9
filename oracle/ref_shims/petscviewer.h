/* see petscsys.h (oracle shim) */
#include <petscsys.h>
