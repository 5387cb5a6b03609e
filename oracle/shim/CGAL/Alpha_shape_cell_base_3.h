/* shim: see wn_cgal_shim.h */
#include "wn_cgal_shim.h"
