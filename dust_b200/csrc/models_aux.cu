// One-step transition of the reference's two remaining forward models.  Neither has a
// cost function or a demo in the reference, so they only exist as `model.step`: a thread per row.
//   skid-steer robot   dust/models/skid_steer_robot.py:73-122  (5 states, 2 wheel speeds; parameters x_icr, wheel_radius,
//                      axial_distance)
//   cart-pole          dust/models/cartpole.py:127-172 (4 states, 1 force; parameters g, mass_cart, mass_pole, length, mu_c,
//                      mu_p, f_mag).  The reference's own step raises AttributeError (it reads the name-mangled
//                      `self.__params_dict`, cartpole.py:150-155); the arithmetic below is the method body as written --
//                      including `mass = m_c + m_c` (sic, :160) -- with that one lookup repaired.
// Compiled with -fmad=false: one rounding per reference operation.
#include <math.h>

#include "common.cuh"

namespace dust {

struct AuxParams {
  int kind, M, np;
  float dt;
  float cfg[8];      // skid: x_icr, wheel_radius, axial_distance, min_r, max_r, min_l, max_l | cart-pole: g, m_c, m_p, length, mu_c, mu_p, f_mag
  const float* states;
  const float* actions;
  const float* params;   // [M, np] per-row overrides in the order above (first 3 / 7 entries), or nullptr
  float* next;
};

__global__ void __launch_bounds__(128) aux_model_step_kernel(const AuxParams k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k.M) return;
  const float* prm = k.params ? k.params + (long long)i * k.np : nullptr;
  const float dt = k.dt;
  if (k.kind == DUST_AUX_SKID_STEER) {
    const float* s = k.states + 5LL * i;
    const float x = s[0], y = s[1], th = s[2];
    const float x_icr = prm ? prm[0] : k.cfg[0], wr = prm ? prm[1] : k.cfg[1], ad = prm ? prm[2] : k.cfg[2];
    const float r = fminf(fmaxf(k.actions[2LL * i], k.cfg[3]), k.cfg[4]);
    const float l = fminf(fmaxf(k.actions[2LL * i + 1], k.cfg[5]), k.cfg[6]);
    const float pi = 3.14159274101257324f;                   // python float math.pi times a float32 tensor
    const float lin = ((r + l) * pi) * wr;                    // skid_steer_robot.py:96
    const float ang = ((((r - l) * 2.0f) * pi) * wr) / ad;    // :97-99
    const float fwd = lin * dt;
    const float lat = ((-ang) * x_icr) * dt;
    const float c = cosf(th), sn = sinf(th);
    float* o = k.next + 5LL * i;
    o[0] = (x + fwd * c) - lat * sn;
    o[1] = (y + fwd * sn) + lat * c;
    o[2] = th + ang * dt;
    o[3] = lin;
    o[4] = ang;
  } else {
    const float* s = k.states + 4LL * i;
    const float x_d = s[1], th = s[2], th_d = s[3];
    const float g = prm ? prm[0] : k.cfg[0], m_c = prm ? prm[1] : k.cfg[1], m_p = prm ? prm[2] : k.cfg[2];
    const float len = prm ? prm[3] : k.cfg[3], mu_c = prm ? prm[4] : k.cfg[4], mu_p = prm ? prm[5] : k.cfg[5];
    const float f_mag = prm ? prm[6] : k.cfg[6];
    const float act = fminf(fmaxf(k.actions[i], -1.0f), 1.0f) * f_mag;
    const float mass = m_c + m_c;                              // sic (cartpole.py:160)
    const float pm = m_p * len;
    const float sgn = (x_d > 0.f) ? 1.0f : ((x_d < 0.f) ? -1.0f : 0.0f);
    const float cart_friction = mu_c * sgn;
    const float pole_friction = (mu_p * th_d) / pm;
    const float sn = sinf(th), c = cosf(th);
    const float factor = ((act + (pm * sn) * (th_d * th_d)) - cart_friction) / mass;
    const float num = (g * sn - c * factor) - pole_friction;
    const float den = len * (1.3333333333333333f - (m_p * (c * c)) / mass);
    const float th_dd = num / den;
    const float x_dd = factor - ((pm * th_dd) * c) / mass;
    float* o = k.next + 4LL * i;
    o[0] = s[0] + x_d * dt;
    o[1] = x_d + x_dd * dt;
    o[2] = th + th_d * dt;
    o[3] = th_d + th_dd * dt;
  }
}

}  // namespace dust

using namespace dust;

extern "C" int dust_aux_model_step(int32_t kind, float dt, const float* cfg, int32_t M, const float* states, const float* actions,
                                   const float* params, float* next_states, void* stream_) {
  DUST_REQUIRE(kind == DUST_AUX_SKID_STEER || kind == DUST_AUX_CARTPOLE, DUST_ERR_UNSUPPORTED, "dust_aux_model_step: unknown model kind %d", kind);
  DUST_REQUIRE(cfg && states && actions && next_states && M > 0 && dt > 0.f, DUST_ERR_INVALID_ARG,
               "dust_aux_model_step: cfg, states, actions, next_states, M > 0 and dt > 0 are required");
  AuxParams k;
  k.kind = kind; k.M = M; k.np = kind == DUST_AUX_SKID_STEER ? 3 : 7; k.dt = dt;
  for (int i = 0; i < 8; ++i) k.cfg[i] = cfg[i];
  k.states = states; k.actions = actions; k.params = params; k.next = next_states;
  cudaStream_t stream = (cudaStream_t)stream_;
  { DUST_TIMED("aux_model_step_kernel", stream); aux_model_step_kernel<<<ceil_div(M, 128), 128, 0, stream>>>(k); }
  DUST_LAUNCH_OK("aux_model_step_kernel");
  return DUST_OK;
}
