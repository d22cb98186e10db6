// Device-resident controller state of one CNF dopri5 solve (shared by the SIMT and the
// tensor-core engines).  Written by the 1-thread controller kernels, read by every other
// kernel of the solve, copied to the host only when the host polls for completion.
#pragma once
#include <stdint.h>

struct CnfState {
  double t;        // current time (start of the step being attempted), torchdiffeq's state.t1
  double t_prev;   // start of the last attempted step (state.t0)
  double dt;       // step size of the step being attempted (float64 like torchdiffeq)
  double t_end;
  double sum_x;    // error-ratio accumulators: sum over the x tensor / the logp tensor
  double sum_l;
  float dt_prev;   // fp32 dt of the last attempted step (dense output)
  float first_dt;
  int32_t status, nfe, accepted, rejected;
  int32_t done;      // solve finished (or failed): every later kernel exits immediately
  int32_t accept;    // decision of the last controller run
  int32_t fin_step;  // id of the step the last controller run belonged to
  int32_t pad;
};
