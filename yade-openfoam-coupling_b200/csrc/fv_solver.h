// fv_solver.h -- internal interface of the finite-volume / PISO half (fv_*.cu).
#pragma once
#include "fy_ctx.h"

void fvDestroy(fy_ctx* h);
