// navsim_math.h — deterministic fp64 math + rigid-body / LiDAR physics shared by the
// sm_100a kernels (navsim_kernels.cu) and by the CPU shim that stands in for Gazebo
// underneath the reference's Env (oracle/).
//
// Why this file exists: the reference quantises yaw to whole degrees, goal offsets to
// 0.1 m and bearing angles to 0.01 deg with Python round() (environment_new.py:142,
// 149-150,169,172-176).  A one-ulp pose difference across such a boundary moves an
// observation feature by up to 0.5, so the GPU simulator and the CPU stand-in for the
// un-vendored Gazebo plugins must produce *bit-identical* poses and ranges.  Everything
// here therefore uses only IEEE-754 correctly rounded primitives (+ - * / sqrt fma rint)
// in a fixed order.  Build rules: device side with `-fmad=false`, host side with
// `-ffp-contract=off`; never with fast-math.
//
// Physics rows:
//   K  differential drive, midpoint form       turtlebot3_fake.cpp:117-118,157-163
//      (v, w) mapping                          environment_new.py:273-278
//      dt = 1 / LiDAR update rate (5 Hz)       turtlebot3_burger.gazebo.xacro:107
//   R  planar ray sensor                       turtlebot3_burger.gazebo.xacro:104-127
//      sensor mount x = -0.032 m               turtlebot3_burger.urdf.xacro:134-138
#ifndef NAVSIM_MATH_H_
#define NAVSIM_MATH_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NV_HD __host__ __device__ __forceinline__
#else
#define NV_HD static inline
#endif

#define NV_INF_F (__builtin_huge_valf())
#define NV_PI        3.141592653589793238462643383279502884
#define NV_TWO_PI    6.283185307179586476925286766559005768
#define NV_RAD2DEG   0x1.ca5dc1a63c1f8p+5 /* 180.0 / pi, as CPython's math.degrees uses */
#define NV_INF       (__builtin_huge_val())

// ---------------------------------------------------------------------------------------
// sin/cos for |x| <= ~8: Cody-Waite reduction by pi/2 (two-term) + degree-13/12 minimax
// kernels on [-pi/4, pi/4].  About 1 ulp; the point is determinism, not the last bit.
// ---------------------------------------------------------------------------------------
NV_HD double nv_ksin(double r) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = r * r;
  double p = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return r + (z * r) * (S1 + z * p);
}

NV_HD double nv_kcos(double r) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = r * r;
  double p = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  return (1.0 - 0.5 * z) + z * p;
}

NV_HD void nv_sincos(double x, double* s, double* c) {
  const double INV_PIO2 = 6.36619772367581382433e-01;
  const double PIO2_HI = 1.57079632673412561417e+00;  // high 33 bits of pi/2
  const double PIO2_LO = 6.07710050650619224932e-11;  // pi/2 - PIO2_HI
  double k = rint(x * INV_PIO2);
  double r = (x - k * PIO2_HI) - k * PIO2_LO;
  double sr = nv_ksin(r), cr = nv_kcos(r);
  int q = ((int)k) & 3;
  double ss = (q & 1) ? cr : sr;
  double cc = (q & 1) ? sr : cr;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}

// ---------------------------------------------------------------------------------------
// atan(x): four-breakpoint reduction + odd minimax polynomial (the classic table of
// atan(0.5), atan(1), atan(1.5), atan(inf) split into hi/lo parts).
// ---------------------------------------------------------------------------------------
NV_HD double nv_atan(double x) {
  const double HI0 = 4.63647609000806093515e-01, LO0 = 2.26987774529616870924e-17;
  const double HI1 = 7.85398163397448278999e-01, LO1 = 3.06161699786838301793e-17;
  const double HI2 = 9.82793723247329054082e-01, LO2 = 1.39033110312309984516e-17;
  const double HI3 = 1.57079632679489655800e+00, LO3 = 6.12323399573676603587e-17;
  const double T0 = 3.33333333333329318027e-01, T1 = -1.99999999998764832476e-01,
               T2 = 1.42857142725034663711e-01, T3 = -1.11111104054623557880e-01,
               T4 = 9.09088713343650656196e-02, T5 = -7.69187620504482999495e-02,
               T6 = 6.66107313738753120669e-02, T7 = -5.83357013379057348645e-02,
               T8 = 4.97687799461593236017e-02, T9 = -3.65315727442169155270e-02,
               T10 = 1.62858201153657823623e-02;
  double ax = fabs(x);
  double hi, lo, t;
  int id;
  if (ax < 0.4375) {
    id = -1; t = ax; hi = 0.0; lo = 0.0;
  } else if (ax < 0.6875) {
    id = 0; t = (2.0 * ax - 1.0) / (2.0 + ax); hi = HI0; lo = LO0;
  } else if (ax < 1.1875) {
    id = 1; t = (ax - 1.0) / (ax + 1.0); hi = HI1; lo = LO1;
  } else if (ax < 2.4375) {
    id = 2; t = (ax - 1.5) / (1.0 + 1.5 * ax); hi = HI2; lo = LO2;
  } else {
    id = 3; t = -1.0 / ax; hi = HI3; lo = LO3;
  }
  double z = t * t, w = z * z;
  double s1 = z * (T0 + w * (T2 + w * (T4 + w * (T6 + w * (T8 + w * T10)))));
  double s2 = w * (T1 + w * (T3 + w * (T5 + w * (T7 + w * T9))));
  double r;
  if (id < 0) r = t - t * (s1 + s2);
  else        r = hi - ((t * (s1 + s2) - lo) - t);
  return (x < 0.0) ? -r : r;
}

// ---------------------------------------------------------------------------------------
// Python round(x, k) for k in {0,1,2}: round-half-even on the EXACT decimal expansion of
// the double, result = the double nearest n / 10^k (float.__round__ goes through
// correctly rounded dtoa/strtod).  x * 10^k is formed exactly as hi + lo with one fma.
// ---------------------------------------------------------------------------------------
NV_HD double nv_round_scaled(double x, double scale) {
  double hi = x * scale;
  double lo = fma(x, scale, -hi);  // exact: x*scale == hi + lo
  double n = rint(hi);             // ties-to-even on hi
  double d = hi - n;               // exact, |d| <= 0.5
  // |d| < 0.5 implies |d| <= 0.5 - ulp(hi) while |lo| <= ulp(hi)/2, so only an exact
  // tie in hi can be overturned by the sign of lo.
  if (d == 0.5 && lo > 0.0) n += 1.0;        // true value just above the tie
  else if (d == -0.5 && lo < 0.0) n -= 1.0;  // true value just below the tie
  return n;
}
// n / scale for an integer-valued n and scale in {10, 100}.  For |n| < 2^26 one Markstein
// correction step on n * RN(1/scale) yields the correctly rounded quotient (checked against
// the division for EVERY such n by tests/test_math_host.py); anything larger divides.  Costs
// three fp64 instructions instead of a division subroutine.
NV_HD double nv_div_scale(double n, double scale, double inv_scale) {
  if (!(fabs(n) < 67108864.0)) return n / scale;
  double q = n * inv_scale;
  double r = fma(-q, scale, n);
  double q1 = fma(r, inv_scale, q);
  return (n == 0.0) ? n : q1;  // keep the sign of a zero
}
NV_HD double nv_pyround0(double x) { return rint(x); }
NV_HD double nv_pyround1(double x) { return nv_div_scale(nv_round_scaled(x, 10.0), 10.0, 0.1); }
NV_HD double nv_pyround2(double x) { return nv_div_scale(nv_round_scaled(x, 100.0), 100.0, 0.01); }

// ---------------------------------------------------------------------------------------
// Env.getOdometry's goal bearing (environment_new.py:149-169) as a function of the two goal
// offsets in tenths of a metre: rel_dis_x = round(goal.x - x, 1) = RN(nx / 10) and likewise
// ny, so rel_theta = round(degrees(theta), 2) depends on the integers (nx, ny) alone.
// Returns rel_theta in hundredths of a degree (0 .. 36000).  The simulator tabulates this
// function once per map (host build of this header) and the step kernel looks it up.
// ---------------------------------------------------------------------------------------
NV_HD int nv_rel_theta_centideg(int nx, int ny) {
  const double rx = nv_div_scale((double)nx, 10.0, 0.1), ry = nv_div_scale((double)ny, 10.0, 0.1);
  double theta;
  if (nx > 0 && ny > 0) theta = nv_atan(ry / rx);                        // :153
  else if (nx > 0 && ny < 0) theta = 2.0 * NV_PI + nv_atan(ry / rx);     // :155
  else if (nx < 0 && ny < 0) theta = NV_PI + nv_atan(ry / rx);           // :157
  else if (nx < 0 && ny > 0) theta = NV_PI + nv_atan(ry / rx);           // :159
  else if (nx == 0 && ny > 0) theta = 0.5 * NV_PI;                       // :161
  else if (nx == 0 && ny < 0) theta = 1.5 * NV_PI;                       // :163
  else if (ny == 0 && nx > 0) theta = 0.0;                               // :165
  else theta = NV_PI;                                                    // :167
  return (int)nv_round_scaled(theta * NV_RAD2DEG, 100.0);                // :169
}

// diff_angle = yaw - rel_theta wrapped to [-180, 180] and rounded to 0.01 (:170-176), in
// hundredths of a degree.  yaw is a whole number of degrees and rel_theta = RN(m / 100), so
// the floating-point difference lies within 1e-13 of the exact hundredth (100 yaw - m) / 100,
// every round(., 2) in :172-176 lands on that hundredth, and the comparisons against 0 and
// +-180 can only tie when the hundredth is the bound itself, where the difference is exact.
NV_HD int nv_diff_angle_centideg(int yaw_deg, int rel_theta_centideg) {
  int k = 100 * yaw_deg - rel_theta_centideg;
  if (k < -18000) k += 36000;       // :174
  else if (k > 18000) k -= 36000;   // :176
  return k;
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG.  Goal sampling draws are keyed (seed, global agent id) and
// counted per agent, so results do not depend on how agents are sharded over GPUs.
// ---------------------------------------------------------------------------------------
NV_HD void nv_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

NV_HD void nv_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                            uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    uint32_t h0, l0, h1, l1;
    nv_mulhilo32(0xD2511F53u, c0, &h0, &l0);
    nv_mulhilo32(0xCD9E8D57u, c2, &h1, &l1);
    uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two 32-bit words, the way CPython's random.random() packs
// its two Mersenne words: (a >> 5) * 2^26 + (b >> 6), scaled by 2^-53.
NV_HD double nv_u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// Draw number `draw` of agent `agent`: one (ux, uy) pair per Philox block.
NV_HD void nv_goal_uniforms(uint64_t seed, uint64_t agent, uint32_t draw, double* ux, double* uy) {
  uint32_t o[4];
  nv_philox4x32_10(draw, 0u, (uint32_t)agent, (uint32_t)(agent >> 32),
                   (uint32_t)seed, (uint32_t)(seed >> 32), o);
  *ux = nv_u53(o[0], o[1]);
  *uy = nv_u53(o[2], o[3]);
}

// Scripted benchmark actions a0 ~ U[0,1), a1 ~ U[-1,1): step s of an agent uses words
// 2 (s & 1), 2 (s & 1) + 1 of the Philox block with counter (s >> 1, 1, agent) keyed by the
// action seed - one block serves two consecutive steps.
NV_HD void nv_scripted_block(uint64_t action_seed, uint64_t agent, uint32_t step, uint32_t o[4]) {
  nv_philox4x32_10(step >> 1, 1u, (uint32_t)agent, (uint32_t)(agent >> 32), (uint32_t)action_seed,
                   (uint32_t)(action_seed >> 32), o);
}
NV_HD void nv_scripted_action(const uint32_t o[4], uint32_t step, float* a0, float* a1) {
  const uint32_t w0 = (step & 1u) ? o[2] : o[0], w1 = (step & 1u) ? o[3] : o[1];
  *a0 = (float)(w0 >> 8) * (1.0f / 16777216.0f);
  *a1 = (float)(w1 >> 8) * (2.0f / 16777216.0f) - 1.0f;
}

// ---------------------------------------------------------------------------------------
// Row K — differential drive over one LiDAR period.  The reference's fake node splits
// (v, w) into wheel speeds and recombines them (turtlebot3_fake.cpp:117-118,150-151);
// algebraically ds = v dt, dth = w dt, which is what is integrated here in midpoint form
// (:157-163).  Heading is kept wrapped to (-pi, pi] so the trig argument stays small.
// ---------------------------------------------------------------------------------------
// Split in two so a caller can evaluate the two sines/cosines a step needs (midpoint heading
// for the motion, new heading for the sensor) side by side: nv_drive_plan gives the
// arguments, nv_drive_apply moves the pose.
NV_HD void nv_drive_plan(double th, double v, double w, double dt, double* ds, double* th_mid, double* th_new) {
  *ds = v * dt;
  double dth = w * dt;
  *th_mid = th + dth / 2.0;
  double t = th + dth;
  if (t > NV_PI) t = t - NV_TWO_PI;
  else if (t <= -NV_PI) t = t + NV_TWO_PI;
  *th_new = t;
}
NV_HD void nv_drive_apply(double* x, double* y, double ds, double s_mid, double c_mid) {
  *x = *x + ds * c_mid;
  *y = *y + ds * s_mid;
}
NV_HD void nv_drive(double* x, double* y, double* th, double v, double w, double dt) {
  double ds, th_mid, th_new, s, c;
  nv_drive_plan(*th, v, w, dt, &ds, &th_mid, &th_new);
  nv_sincos(th_mid, &s, &c);
  nv_drive_apply(x, y, ds, s, c);
  *th = th_new;
}

// Fidelity option: the diff-drive plugin's wheel acceleration limit
// (turtlebot3_burger.gazebo.xacro:67 <wheelAcceleration>, update rate 30 Hz :71).  The plugin's source is
// not vendored; the model stated here is the one its parameters describe: at every plugin update (`substeps`
// per LiDAR period) each wheel's rim speed moves toward its target by at most accel * dt_sub, and the pose is
// integrated in the same midpoint form with the speeds of that sub-step.  vl / vr are the rim speeds in m/s
// (state carried across steps, zero after a reset); targets from (v, w) as turtlebot3_fake.cpp:117-118.
// accel <= 0 must not reach this function (the caller takes the plain nv_drive path).
NV_HD void nv_drive_ramped(double* x, double* y, double* th, double* vl, double* vr, double v, double w, double dt,
                           double accel, double wheel_sep, int substeps) {
  const double tl = v - w * wheel_sep / 2.0, tr = v + w * wheel_sep / 2.0;
  const double h = dt / (double)substeps, dmax = accel * h;
  for (int k = 0; k < substeps; ++k) {
    double dl = tl - *vl, dr = tr - *vr;
    dl = dl > dmax ? dmax : (dl < -dmax ? -dmax : dl);
    dr = dr > dmax ? dmax : (dr < -dmax ? -dmax : dr);
    *vl = *vl + dl;
    *vr = *vr + dr;
    nv_drive(x, y, th, (*vr + *vl) / 2.0, (*vr - *vl) / wheel_sep, h);
  }
}

// GoalSpawnSampler.sample_start_and_goal (spawn_goal_sampler.py:52-63): draw a start pose and a goal point
// from the two tables until their distance lies in [min_dist, max_dist], at most 100 times, then take one
// more pair unconditionally.  The reference draws rng.randint(len(table)) from numpy's Mersenne Twister;
// here draw k of an agent is Philox block (k, 3, agent) and an index is floor(u32 * n / 2^32).
NV_HD void nv_table_indices(uint64_t seed, uint64_t agent, uint32_t draw, int n_starts, int n_goals, int* is, int* ig) {
  uint32_t o[4], hi, lo;
  nv_philox4x32_10(draw, 3u, (uint32_t)agent, (uint32_t)(agent >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  nv_mulhilo32(o[0], (uint32_t)n_starts, &hi, &lo);
  *is = (int)hi;
  nv_mulhilo32(o[1], (uint32_t)n_goals, &hi, &lo);
  *ig = (int)hi;
}
NV_HD void nv_sample_tables(uint64_t seed, uint64_t agent, uint32_t* draws, const double* starts, int n_starts,
                            const double* goals, int n_goals, double min_dist, double max_dist, int* is_out, int* ig_out) {
  int is = 0, ig = 0;
  for (int attempt = 0; attempt < 100; ++attempt) {                     // spawn_goal_sampler.py:53-54
    nv_table_indices(seed, agent, *draws, n_starts, n_goals, &is, &ig);
    *draws += 1u;
    const double dx = starts[3 * is] - goals[2 * ig], dy = starts[3 * is + 1] - goals[2 * ig + 1];
    const double dist = sqrt(dx * dx + dy * dy);                        // :57 np.linalg.norm
    if (min_dist <= dist && dist <= max_dist) { *is_out = is; *ig_out = ig; return; }   // :58-59
  }
  nv_table_indices(seed, agent, *draws, n_starts, n_goals, &is, &ig);   // :60-62
  *draws += 1u;
  *is_out = is; *ig_out = ig;
}

// Fidelity option: Gaussian range noise of the ray sensor (turtlebot3_burger.gazebo.xacro:122-126, stddev
// 0.01).  Four standard normals for beams 4 b4 .. 4 b4 + 3 of one scan: Philox block (b4, 4, agent) keyed by
// the seed and mixed with the scan's identity (episode draw counter, step), Box-Muller in fp32.
NV_HD void nv_scan_noise4(uint64_t seed, uint64_t agent, uint32_t episode_draws, uint32_t step, uint32_t b4, float n[4]) {
  uint32_t o[4];
  nv_philox4x32_10(b4 | (step << 8), 4u + (episode_draws << 4), (uint32_t)agent, (uint32_t)(agent >> 32),
                   (uint32_t)seed, (uint32_t)(seed >> 32), o);
  for (int k = 0; k < 2; ++k) {
    const float u = ((float)(o[2 * k] >> 8) + 0.5f) * (1.0f / 16777216.0f);       // (0, 1)
    const float v = ((float)(o[2 * k + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u));
    n[2 * k] = r * cosf(6.283185307179586f * v);
    n[2 * k + 1] = r * sinf(6.283185307179586f * v);
  }
}
// the noisy range re-enters the sensor's gates (a hit pushed past max reads +inf, below min -inf)
NV_HD float nv_noisy_range(float t, float noise, float sigma, float rmin, float rmax) {
  if (!(t < NV_INF_F) || !(t > -NV_INF_F)) return t;   // no hit / already gated
  const float r = t + sigma * noise;
  if (r > rmax) return NV_INF_F;
  if (r < rmin) return -NV_INF_F;
  return r;
}

// ---------------------------------------------------------------------------------------
// Row R — planar ray sensor, in fp32 like the float32 ranges a LaserScan carries (the
// sensor's own resolution is 15 mm, gazebo.xacro:120).  The robot pose stays fp64; only
// the sensor origin, the beam directions and the wall segments are rounded to fp32.
//
// A wall segment p0 -> p1 is packed as 8 floats
//     {x0, y0, ex, ey,  mx, my, R2, lim}      e = p1 - p0, m = midpoint,
//     R2 = (rmax + |e|/2)^2 (bounding-circle reach), lim = rmax * |e| (line reach).
// For a ray o + t d:  w = p0 - o,  tn = w x e,  den = d x e,  un = w x d,
//                     t = tn / den,  u = un / den, hit iff t >= 0 and 0 <= u <= 1.
// tn (= |e| * signed distance from o to the wall's line) does not depend on the beam, so
// the per-segment setup is done once for all beams:
//   * range cull   — |tn| > lim or |m - o|^2 > R2: every hit would lie beyond rmax and be
//                    gated to +inf anyway, so skipping the wall cannot change a scan;
//   * back faces   — for maps made of closed convex boxes (edges counter-clockwise) a wall
//                    seen from behind (tn >= 0) is always preceded by a front face;
//   * orientation  — (w, e) are flipped so tn > 0 and a hit needs den > 0, and 1/tn is
//                    formed once, so a beam tracks the nearest hit as the LARGEST inverse
//                    distance q = den / tn with one multiply and one max per test.
// u is accepted on [-eps, 1 + eps] so a ray through the shared corner of two box edges
// cannot slip between them.  All products that feed a difference are explicit fmaf so the
// host and device builds round identically whatever the contraction flags.
// ---------------------------------------------------------------------------------------
#define NV_SEG_FLOATS 8
#define NV_SEG_EPS 1.0e-6f
#define NV_MAP_CLOSED_BOXES 1

// {x0,y0,x1,y1} metres (double) -> packed fp32 record.
NV_HD void nv_pack_segment(const double* xyxy, double rmax, float* out) {
  double ex = xyxy[2] - xyxy[0], ey = xyxy[3] - xyxy[1];
  double len = sqrt(ex * ex + ey * ey);
  double reach = rmax * 1.001 + 0.5 * len + 1.0e-3;   // conservative by construction
  out[0] = (float)xyxy[0];
  out[1] = (float)xyxy[1];
  out[2] = (float)ex;
  out[3] = (float)ey;
  out[4] = (float)(0.5 * (xyxy[0] + xyxy[2]));
  out[5] = (float)(0.5 * (xyxy[1] + xyxy[3]));
  out[6] = (float)(reach * reach);
  out[7] = (float)((rmax * 1.001 + 1.0e-3) * len);
}

typedef struct nv_seg_view { float wx, wy, ex, ey, inv_tn; } nv_seg_view;

// Per-segment setup for a sensor origin.  Returns 0 when no beam can see the wall.  The three
// culls are independent predicates of (wall, origin), so their order is free: the cheapest and
// most selective come first (half of a box's edges face away; most others are out of range).
// nv_seg_cull is the visibility test alone (the step kernel uses it to compact the walls an
// agent can see before it casts any beam); nv_seg_setup adds the reciprocal a beam test needs.
NV_HD int nv_seg_cull(const float* sg, float ox, float oy, int closed_boxes, float* wx_o, float* wy_o, float* ex_o,
                      float* ey_o, float* tn_o) {
  float wx = sg[0] - ox, wy = sg[1] - oy, ex = sg[2], ey = sg[3];
  float tn = fmaf(wx, ey, -(wy * ex));
  if (closed_boxes && tn >= 0.0f) return 0;
  if (tn < 0.0f) { tn = -tn; wx = -wx; wy = -wy; ex = -ex; ey = -ey; }
  if (tn > sg[7] || tn == 0.0f) return 0;
  float cx = sg[4] - ox, cy = sg[5] - oy;
  if (fmaf(cx, cx, cy * cy) > sg[6]) return 0;
  *wx_o = wx; *wy_o = wy; *ex_o = ex; *ey_o = ey; *tn_o = tn;
  return 1;
}

NV_HD int nv_seg_setup(const float* sg, float ox, float oy, int closed_boxes, nv_seg_view* v) {
  float tn;
  if (!nv_seg_cull(sg, ox, oy, closed_boxes, &v->wx, &v->wy, &v->ex, &v->ey, &tn)) return 0;
  v->inv_tn = 1.0f / tn;
  return 1;
}

// One beam against one prepared wall: returns max(best_q, q) when the beam hits it.
NV_HD float nv_ray_q(const nv_seg_view* v, float dx, float dy, float best_q) {
  float den = fmaf(dx, v->ey, -(dy * v->ex));
  float un = fmaf(v->wx, dy, -(v->wy * dx));
  float q = den * v->inv_tn;
  int ok = (fmaf(den, NV_SEG_EPS, un) >= 0.0f) && (fmaf(den, -NV_SEG_EPS, un) <= den) && (q > best_q);
  return ok ? q : best_q;
}

// One beam against all S walls (beam-major form): inverse distance of the nearest hit.
NV_HD float nv_beam_q(float ox, float oy, float dx, float dy, const float* seg, int S, int closed_boxes) {
  float best_q = 0.0f;
  for (int k = 0; k < S; ++k) {
    nv_seg_view v;
    if (nv_seg_setup(seg + NV_SEG_FLOATS * k, ox, oy, closed_boxes, &v)) best_q = nv_ray_q(&v, dx, dy, best_q);
  }
  return best_q;
}

// Beam direction = heading rotated by the beam's table entry (cos, sin of its offset).
NV_HD void nv_beam_dir(float ch, float sh, float bc, float bs, float* dx, float* dy) {
  *dx = fmaf(ch, bc, -(sh * bs));
  *dy = fmaf(sh, bc, ch * bs);
}

// Inverse distance -> range with the gates of the Gazebo ray sensor: no hit or beyond
// max -> +inf, below min -> -inf (gazebo.xacro:117-119).
NV_HD float nv_range_from_q(float q, float rmin, float rmax) {
  float t = 1.0f / q;  // q == 0 (no hit) gives +inf
  if (t > rmax) return NV_INF_F;
  if (t < rmin) return -NV_INF_F;
  return t;
}

#endif  // NAVSIM_MATH_H_
