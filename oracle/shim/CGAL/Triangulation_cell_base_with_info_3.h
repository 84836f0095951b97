// TEST INFRASTRUCTURE ONLY: see flat_triangulation_3.h
#include "flat_triangulation_3.h"
