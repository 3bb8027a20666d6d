/*
 * sag_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See sag_oracle.h.
 *
 * Plain C99, float64, sequential, one environment at a time.  Every block cites the
 * reference lines (relative to /root/reference/) it restates.  Compile with
 * -ffp-contract=off so that the arithmetic is plain IEEE double.
 */
#include "sag_oracle.h"
#include "../include/sag_detmath.h" /* deterministic sin/cos/atan2/log (bit-identical on the GPU) */

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846
#define TWO_PI (2.0 * PI)
#define GRAV 9.81

/* ------------------------------------------------------------------------------------------
 * Model constants (derived from assets/xmls/point.xml:2-38 with MuJoCo's
 * uniform-density rule [EXT])
 * ---------------------------------------------------------------------------------------- */
#define PT_TIMESTEP 0.004      /* point.xml:3 */
#define PT_NSUB 5              /* safe_adaptation_gym.py:17 */
#define PT_R 0.1               /* point.xml:18 sphere size */
#define PT_ARROW_OFF 0.1       /* point.xml:19 pos */
#define PT_ARROW_H 0.05        /* point.xml:19 size */
#define PT_FORCE_LIM 0.05      /* point.xml:7-8 forcerange */
#define PT_GEAR_X 0.3          /* point.xml:36 */
#define PT_GEAR_Z 0.3          /* point.xml:37 */
#define PT_DAMP_XY 0.01        /* point.xml:15-16 */
#define PT_DAMP_Z 0.005        /* point.xml:17 */
#define PT_Z 0.1               /* point.xml:13 */
#define GOAL_Z 0.16            /* primitive_objects.py:140: size/2 + 1e-2 with size 0.3 */
#define GOAL_SIZE 0.3          /* go_to_goal.py:12 */
#define GOAL_KEEPOUT 0.4       /* go_to_goal.py:13 */
#define BUTTON_SIZE 0.1        /* press_buttons.py:15 */
#define BUTTONS_KEEPOUT 0.2    /* press_buttons.py:14 */
#define BUTTON_DELAY 5         /* press_buttons.py:17 */
#define BOX_SIZE 0.2           /* push_box.py:12 */
#define BOX_DENSITY 0.001      /* push_box.py:15 */
#define VASES_DENSITY 0.001    /* consts.py:21 */
#define BALL_R 0.14            /* dribble_ball.py:11 */
#define BALL_DENSITY (BOX_DENSITY / 2.0) /* dribble_ball.py:31 */
#define BALL_SOL_TC 0.018      /* dribble_ball.py:34 solref = (0.018, 0.2) */
#define BALL_SOL_DR 0.2
#define BALL_ROLL 0.05         /* dribble_ball.py:32 friction = (slide 1.2, spin 0.003, roll 0.05); condim 6 */
#define BALL_SPIN 0.003
#define ROD_R 0.08             /* roll_rod.py:12 */
#define ROD_HALF 0.3           /* roll_rod.py:11: cylinder size = (radius, half length) */
#define ROD_DENSITY (BOX_DENSITY / 2.0)  /* roll_rod.py:35 */
#define ROD_ROLL 0.05          /* roll_rod.py:36 friction = (1.2, 0.001, 0.05); floor condim 6 (point.xml:12) */
#define PRIO_MU 1.2            /* ball / rod geoms have priority 1: their sliding friction wins the mix [EXT] */
#define PI_D 3.14159265358979323846
#define LIDAR_MAX_DIST 5.0     /* safe_adaptation_gym.py:23 */
#define TENDON_MAX (BOX_SIZE * 3.75) /* haul_box.py:25 */

/* MuJoCo default soft-constraint parameters [EXT]: solref (0.02, 1), solimp (0.9,0.95,0.001,0.5,2) */
#define SOL_TC 0.02
#define SOL_DR 1.0
#define WELD_SOL_TC 0.02   /* primitive_objects.py:80-82: <weld solref=".02 1.5"/> */
#define WELD_SOL_DR 1.5
#define IMP_D0 0.9
#define IMP_DMAX 0.95
#define IMP_WIDTH 0.001
#define FRICTION_MU 1.0        /* max(geom frictions) = 1 for every pair on this path */
#define PGS_SWEEPS 10
#define PGS_TOL 1e-8
#define SLEEP_V 1e-8

/* car.xml: timestep 0.008, frame_skip 10 (safe_adaptation_gym.py:18); default geom density 5 (car.xml:5) */
#define CAR_TIMESTEP 0.008
#define CAR_NSUB 10
#define CAR_FORCE_LIM 0.02     /* car.xml:7 */
#define CAR_WHEEL_R 0.05       /* car.xml:5 default size; fromto cylinders */
#define CAR_WHEEL_DAMP 0.001   /* car.xml:6 */
#define CAR_ARMATURE 0.00025   /* car.xml:22,26 */
#define CAR_DENSITY 5.0
#define CAR_NGEOM 8
/* robot geoms in the body frame (x, y, half-x, half-y | radius), car.xml:16-31; wheels: x-axis cylinders -> footprints */
static const double CAR_GEOM[CAR_NGEOM][5] = {
    {0.0, 0.0, 0.1, 0.1, 0.0},       /* robot */
    {0.0, 0.15, 0.1, 0.01, 0.0},     /* back_bumper */
    {0.0, 0.125, 0.01, 0.025, 0.0},  /* back_connector */
    {0.0, -0.165, 0.05, 0.01, 0.0},  /* front_bumper */
    {0.0, -0.13, 0.05, 0.03, 0.0},   /* front_connector */
    {-0.13, 0.1, 0.025, 0.05, 0.0},  /* left wheel footprint */
    {0.13, 0.1, 0.025, 0.05, 0.0},   /* right wheel footprint */
    {0.0, -0.1, 0.0, 0.0, 0.05},     /* rear castor sphere */
};
static const double CAR_GEOM_HZ[5] = {0.05, 0.05, 0.03, 0.05, 0.01}; /* box half heights (masses) */
typedef struct { double M, mcx, mcy, Io, Iw, nwheel, nrear; } car_model;
static car_model car_params(void) {
  car_model c;
  double M = 0.0, mx = 0.0, my = 0.0, Io = 0.0;
  for (int g = 0; g < 5; ++g) { /* boxes */
    double hx = CAR_GEOM[g][2], hy = CAR_GEOM[g][3], hz = CAR_GEOM_HZ[g];
    double m = 8.0 * hx * hy * hz * CAR_DENSITY;
    double x = CAR_GEOM[g][0], y = CAR_GEOM[g][1];
    M += m; mx += m * x; my += m * y;
    Io += m * (4.0 * hx * hx + 4.0 * hy * hy) / 12.0 + m * (x * x + y * y);
  }
  double mw = PI * CAR_WHEEL_R * CAR_WHEEL_R * 0.05 * CAR_DENSITY; /* cylinder radius .05 length .05 */
  for (int g = 5; g < 7; ++g) {
    double x = CAR_GEOM[g][0], y = CAR_GEOM[g][1];
    M += mw; mx += mw * x; my += mw * y;
    Io += mw * (3.0 * CAR_WHEEL_R * CAR_WHEEL_R + 0.05 * 0.05) / 12.0 + mw * (x * x + y * y);
  }
  double mb = 4.0 / 3.0 * PI * CAR_WHEEL_R * CAR_WHEEL_R * CAR_WHEEL_R * CAR_DENSITY;
  { double x = CAR_GEOM[7][0], y = CAR_GEOM[7][1];
    M += mb; mx += mb * x; my += mb * y;
    Io += 0.4 * mb * CAR_WHEEL_R * CAR_WHEEL_R + mb * (x * x + y * y); }
  c.M = M; c.mcx = mx; c.mcy = my; c.Io = Io;
  c.Iw = 0.5 * mw * CAR_WHEEL_R * CAR_WHEEL_R + CAR_ARMATURE;
  /* static normal loads on a level floor: wheels at y = +0.1, castor at y = -0.1 */
  double yc = my / M;
  c.nrear = M * GRAV * (0.1 - yc) / 0.2;
  c.nwheel = (M * GRAV - c.nrear) / 2.0;
  return c;
}

static double pt_mass(void) { return 4.0 / 3.0 * PI * PT_R * PT_R * PT_R + 8.0 * PT_ARROW_H * PT_ARROW_H * PT_ARROW_H; }
static double pt_mc(void) { return 8.0 * PT_ARROW_H * PT_ARROW_H * PT_ARROW_H * PT_ARROW_OFF; } /* m * c */
static double pt_inertia_o(void) {
  double ms = 4.0 / 3.0 * PI * PT_R * PT_R * PT_R, ma = 8.0 * PT_ARROW_H * PT_ARROW_H * PT_ARROW_H;
  double a = 2.0 * PT_ARROW_H;
  return 0.4 * ms * PT_R * PT_R + ma * (a * a + a * a) / 12.0 + ma * PT_ARROW_OFF * PT_ARROW_OFF;
}

/* ------------------------------------------------------------------------------------------ */
struct orc_env {
  int robot, task;
  orc_config cfg;
  /* physics (stand-in for mujoco.Physics built by mujoco_bridge.py:41-168) */
  double h;
  int nsub;
  double damp_x, damp_y, damp_z, gear_x, gear_z;
  double ctrl[2], ctrl_lo[2], ctrl_hi[2];
  double q[3], v[3], qacc[3];
  double wheel_w[2], wheel_tau[2]; /* car: wheel spin rates, constraint torque of the last forward pass */
  double cq[4];                    /* car: castor ball orientation relative to the chassis (ball joint quaternion) */
  double time;
  int nobj;
  orc_obj obj[ORC_MAX_OBJ];
  double oacc[ORC_MAX_OBJ][3];
  int ncon;
  orc_contact con[ORC_MAX_CON];
  int error;
  int tendon_slot; /* haul_box.py:21-30; -1 = none */
  /* gremlins (primitive_objects.py:57-86): weld between the free body and its mocap body.  The mocap bodies sit at the
   * world origin in the XML (the geom carries the offset), so the weld's relative pose at compile time is the
   * gremlin's spawn pose and the weld target is spawn pose + mocap_pos [EXT].  mocap_pos: what set_mocap_pos wrote;
   * mocap_kin: what the last kinematics pass saw (stale for the first substep of physics.step) */
  double mocap_pos[2], mocap_kin[2];
  double gspawn[ORC_MAX_OBJ][3];
  int touched[ORC_MAX_OBJ]; /* movable body had an active constraint row in the last forward pass */
  int overflow;             /* contact / body limit exceeded in the last forward pass: constraint solve skipped */
  /* world / task (world.py, tasks/ *.py) */
  double extents[4];
  int has_rect[ORC_MAX_OBJ];
  double rect[ORC_MAX_OBJ][4];
  double robot_rot, bound;
  int new_task;            /* the next reset builds the World of a fresh Task instance (world.py:72-78) */
  double last_dist[2];
  int goal_button, btn_state, btn_timer;
  unsigned active_mask;
  double cg_cur, cg_next, cg_ox, cg_oy;
  int cg_timer;
  int goal_slot, box_slot, first_button, nbuttons;
  /* rng */
  int replay_mode;
  uint64_t seed;
  uint32_t gid, episode, ctr[3];
  const double* replay;
  int rn, rpos;
  long draws_left;
};

/* ==========================================================================================
 * Philox4x32-10 (Salmon et al., SC'11) -- counter-based RNG shared (by specification) with the
 * GPU.  Known answers are checked in tests/test_oracle_kat.py.
 * ======================================================================================== */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* two uniforms in [0,1) with 53 random bits each, numpy's random_sample construction */
void orc_philox_uniform2(uint64_t seed, uint32_t ctr, uint32_t episode, uint32_t gid, uint32_t stream, double* u2) {
  uint32_t c[4] = {ctr, episode, gid, stream}, k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
  orc_philox4x32_10(c, k, r);
  u2[0] = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6)) * (1.0 / 9007199254740992.0);
  u2[1] = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6)) * (1.0 / 9007199254740992.0);
}

static void rng_pair(orc_env* e, int stream, double* u1, double* u2) {
  if (e->replay_mode) {
    *u1 = (e->rpos < e->rn) ? e->replay[e->rpos] : 0.5; e->rpos++;
    *u2 = (e->rpos < e->rn) ? e->replay[e->rpos] : 0.5; e->rpos++;
    return;
  }
  double u[2];
  orc_philox_uniform2(e->seed, e->ctr[stream]++, e->episode, e->gid, (uint32_t)stream, u);
  *u1 = u[0]; *u2 = u[1];
}
static double rng_single(orc_env* e, int stream) {
  if (e->replay_mode) {
    double u = (e->rpos < e->rn) ? e->replay[e->rpos] : 0.5; e->rpos++;
    return u;
  }
  double u[2];
  orc_philox_uniform2(e->seed, e->ctr[stream]++, e->episode, e->gid, (uint32_t)stream, u);
  return u[0];
}

/* ==========================================================================================
 * Planar collision primitives (restating the MuJoCo contact convention [EXT]: dist = signed
 * distance, normal from geom1 to geom2, pos = midpoint between the two surfaces; a contact is
 * listed when dist <= margin (= 0)).
 * ======================================================================================== */
typedef struct { double cx, cy, c, s, hx, hy; } obox;

int orc_collide_circle_circle(double ax, double ay, double ra, double bx, double by, double rb, double* o) {
  double dx = bx - ax, dy = by - ay;
  double len = sqrt(dx * dx + dy * dy);
  double dist = len - ra - rb;
  if (dist > 0.0) return 0;
  double nx = 1.0, ny = 0.0;
  if (len > 1e-14) { nx = dx / len; ny = dy / len; }
  o[0] = nx; o[1] = ny;
  o[2] = ax + nx * (ra + 0.5 * dist);
  o[3] = ay + ny * (ra + 0.5 * dist);
  o[4] = dist;
  return 1;
}

static int circle_box(double cx, double cy, double r, const obox* B, int circle_is_a, double* o) {
  double rx = cx - B->cx, ry = cy - B->cy;
  double lx = rx * B->c + ry * B->s, ly = -rx * B->s + ry * B->c;
  double qx = lx < -B->hx ? -B->hx : (lx > B->hx ? B->hx : lx);
  double qy = ly < -B->hy ? -B->hy : (ly > B->hy ? B->hy : ly);
  double nlx, nly, dist;
  if (qx == lx && qy == ly) { /* centre inside the box: exit through the nearest face */
    double penx = B->hx - fabs(lx), peny = B->hy - fabs(ly);
    if (penx <= peny) { nlx = lx >= 0.0 ? 1.0 : -1.0; nly = 0.0; qx = nlx * B->hx; dist = -penx - r; }
    else { nlx = 0.0; nly = ly >= 0.0 ? 1.0 : -1.0; qy = nly * B->hy; dist = -peny - r; }
  } else {
    double ex = lx - qx, ey = ly - qy;
    double len = sqrt(ex * ex + ey * ey);
    dist = len - r;
    if (dist > 0.0) return 0;
    nlx = ex / len; nly = ey / len;
  }
  /* world: normal box -> circle, box surface point, circle surface point */
  double nwx = nlx * B->c - nly * B->s, nwy = nlx * B->s + nly * B->c;
  double pbx = B->cx + qx * B->c - qy * B->s, pby = B->cy + qx * B->s + qy * B->c;
  double pcx = cx - nwx * r, pcy = cy - nwy * r;
  o[2] = 0.5 * (pbx + pcx); o[3] = 0.5 * (pby + pcy); o[4] = dist;
  if (circle_is_a) { o[0] = -nwx; o[1] = -nwy; } else { o[0] = nwx; o[1] = nwy; }
  return 1;
}

int orc_collide_circle_box(double cx, double cy, double r, double bx, double by, double byaw, double hx, double hy,
                           int circle_is_a, double* out5) {
  obox B = {bx, by, sag_cos(byaw), sag_sin(byaw), hx, hy};
  return circle_box(cx, cy, r, &B, circle_is_a, out5);
}

/* 2-D SAT + reference-face clipping (1 or 2 points). */
static int box_box(const obox* A, const obox* B, double* o /* [2][5] */) {
  double dx = B->cx - A->cx, dy = B->cy - A->cy;
  double ax[4][2] = {{A->c, A->s}, {-A->s, A->c}, {B->c, B->s}, {-B->s, B->c}};
  double best = -1e300, bsign = 1.0;
  int bi = -1;
  for (int k = 0; k < 4; ++k) {
    double ux = ax[k][0], uy = ax[k][1];
    double dp = dx * ux + dy * uy;
    double ra = A->hx * fabs(A->c * ux + A->s * uy) + A->hy * fabs(-A->s * ux + A->c * uy);
    double rb = B->hx * fabs(B->c * ux + B->s * uy) + B->hy * fabs(-B->s * ux + B->c * uy);
    double sep = fabs(dp) - (ra + rb);
    if (sep > 0.0) return 0;
    if (sep > best) { best = sep; bi = k; bsign = dp >= 0.0 ? 1.0 : -1.0; }
  }
  const obox *R, *I;
  double Nx, Ny; /* reference-face normal, pointing from the reference box to the incident box */
  int ref_is_a = bi < 2;
  if (ref_is_a) { R = A; I = B; Nx = bsign * ax[bi][0]; Ny = bsign * ax[bi][1]; }
  else { R = B; I = A; Nx = -bsign * ax[bi][0]; Ny = -bsign * ax[bi][1]; }
  int ref_axis = bi & 1; /* 0: normal along the reference box's local x, 1: local y */
  double hn = ref_axis == 0 ? R->hx : R->hy, ht = ref_axis == 0 ? R->hy : R->hx;
  double Tx = -Ny, Ty = Nx;
  /* incident face: the face of I whose outward normal is most anti-parallel to N */
  double nlx = Nx * I->c + Ny * I->s, nly = -Nx * I->s + Ny * I->c;
  double v1x, v1y, v2x, v2y;
  if (fabs(nlx) >= fabs(nly)) { double sx = nlx > 0.0 ? -1.0 : 1.0; v1x = sx * I->hx; v1y = -I->hy; v2x = sx * I->hx; v2y = I->hy; }
  else { double sy = nly > 0.0 ? -1.0 : 1.0; v1x = -I->hx; v1y = sy * I->hy; v2x = I->hx; v2y = sy * I->hy; }
  double w1x = I->cx + v1x * I->c - v1y * I->s - R->cx, w1y = I->cy + v1x * I->s + v1y * I->c - R->cy;
  double w2x = I->cx + v2x * I->c - v2y * I->s - R->cx, w2y = I->cy + v2x * I->s + v2y * I->c - R->cy;
  double n1 = w1x * Nx + w1y * Ny, t1 = w1x * Tx + w1y * Ty;
  double n2 = w2x * Nx + w2y * Ny, t2 = w2x * Tx + w2y * Ty;
  double lo = 0.0, hi = 1.0, dt = t2 - t1;
  if (t1 > ht && t2 > ht) return 0;
  if (t1 < -ht && t2 < -ht) return 0;
  if (dt != 0.0) {
    if (t1 > ht) { double s = (ht - t1) / dt; if (s > lo) lo = s; }
    if (t2 > ht) { double s = (ht - t1) / dt; if (s < hi) hi = s; }
    if (t1 < -ht) { double s = (-ht - t1) / dt; if (s > lo) lo = s; }
    if (t2 < -ht) { double s = (-ht - t1) / dt; if (s < hi) hi = s; }
  }
  if (lo > hi) return 0;
  int n = 0;
  double ts[2] = {lo, hi};
  int npts = (hi - lo) > 1e-12 ? 2 : 1;
  for (int k = 0; k < npts; ++k) {
    double s = ts[k];
    double pn = n1 + s * (n2 - n1), pt = t1 + s * dt;
    double sep = pn - hn;
    if (sep > 0.0) continue;
    /* world point on the incident edge, moved half-way back to the reference face */
    double px = R->cx + pn * Nx + pt * Tx - 0.5 * sep * Nx;
    double py = R->cy + pn * Ny + pt * Ty - 0.5 * sep * Ny;
    double* c = o + 5 * n;
    if (ref_is_a) { c[0] = Nx; c[1] = Ny; } else { c[0] = -Nx; c[1] = -Ny; }
    c[2] = px; c[3] = py; c[4] = sep;
    ++n;
  }
  return n;
}

int orc_collide_box_box(double ax, double ay, double ayaw, double ahx, double ahy, double bx, double by, double byaw,
                        double bhx, double bhy, double* out10) {
  obox A = {ax, ay, sag_cos(ayaw), sag_sin(ayaw), ahx, ahy}, B = {bx, by, sag_cos(byaw), sag_sin(byaw), bhx, bhy};
  return box_box(&A, &B, out10);
}

/* ------------------------------------------------------------------------------------------
 * Geoms of each body (primitive_objects.py:39-169, push_box.py:28-72, point.xml:18-19)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int is_box; double cx, cy, c, s, hx, hy, r; } geom2;

static int obj_collidable(int type) {
  return type == ORC_VASE || type == ORC_GREMLIN || type == ORC_PILLAR || type == ORC_BUTTON || type == ORC_BOX ||
         type == ORC_ROD || type == ORC_BALL;
}
static int obj_movable(int type) {
  return type == ORC_VASE || type == ORC_GREMLIN || type == ORC_BOX || type == ORC_ROD || type == ORC_BALL;
}
static int obj_nparts(int type) { return type == ORC_BOX ? 5 : (obj_collidable(type) ? 1 : 0); }

static void obj_geom(const orc_env* e, int slot, int part, geom2* g) {
  const orc_obj* o = &e->obj[slot];
  double c = sag_cos(o->yaw), s = sag_sin(o->yaw);
  g->c = c; g->s = s; g->cx = o->x; g->cy = o->y; g->r = 0.0; g->hx = g->hy = 0.0;
  switch (o->type) {
    case ORC_VASE: g->is_box = 1; g->hx = g->hy = e->cfg.vases_size; break;        /* primitive_objects.py:46-47 */
    case ORC_GREMLIN: g->is_box = 1; g->hx = g->hy = e->cfg.gremlins_size; break;  /* :63-66 */
    case ORC_PILLAR: g->is_box = 0; g->r = e->cfg.pillars_size; break;             /* :117-118 */
    case ORC_BUTTON: g->is_box = 0; g->r = BUTTON_SIZE; break;                     /* :158-161 */
    case ORC_BOX: {                                                                 /* push_box.py:36-67 */
      g->is_box = 1;
      if (part == 0) { g->hx = g->hy = BOX_SIZE; }
      else {
        static const double sx[5] = {0, 1, -1, 1, -1}, sy[5] = {0, 1, 1, -1, -1};
        double ox = sx[part] * BOX_SIZE, oy = sy[part] * BOX_SIZE;
        g->hx = g->hy = BOX_SIZE / 2;
        g->cx = o->x + ox * c - oy * s; g->cy = o->y + ox * s + oy * c;
      }
    } break;
    /* planar footprints: the sphere's great circle (dribble_ball.py:24-30); the lying cylinder's rectangle, axis along
     * body y after euler = (90, 0, 0) (roll_rod.py:27-34) */
    case ORC_BALL: g->is_box = 0; g->r = BALL_R; break;
    case ORC_ROD: g->is_box = 1; g->hx = ROD_R; g->hy = ROD_HALF; break;
    default: g->is_box = 0; break;
  }
}
static int robot_nparts(const orc_env* e) { return e->robot == ORC_CAR ? CAR_NGEOM : 2; }
static void robot_geom(const orc_env* e, int part, geom2* g) {
  double c = sag_cos(e->q[2]), s = sag_sin(e->q[2]);
  g->c = c; g->s = s;
  if (e->robot == ORC_CAR) {
    const double* G = CAR_GEOM[part];
    g->cx = e->q[0] + G[0] * c - G[1] * s; g->cy = e->q[1] + G[0] * s + G[1] * c;
    g->is_box = G[4] == 0.0; g->hx = G[2]; g->hy = G[3]; g->r = G[4];
    return;
  }
  if (part == 0) { g->is_box = 0; g->cx = e->q[0]; g->cy = e->q[1]; g->r = PT_R; g->hx = g->hy = 0; }
  else { g->is_box = 1; g->cx = e->q[0] + PT_ARROW_OFF * c; g->cy = e->q[1] + PT_ARROW_OFF * s; g->hx = g->hy = PT_ARROW_H; g->r = 0; }
}
static int collide(const geom2* A, const geom2* B, double* o) {
  if (!A->is_box && !B->is_box) return orc_collide_circle_circle(A->cx, A->cy, A->r, B->cx, B->cy, B->r, o);
  if (!A->is_box) { obox b = {B->cx, B->cy, B->c, B->s, B->hx, B->hy}; return circle_box(A->cx, A->cy, A->r, &b, 1, o); }
  if (!B->is_box) { obox a = {A->cx, A->cy, A->c, A->s, A->hx, A->hy}; return circle_box(B->cx, B->cy, B->r, &a, 0, o); }
  obox a = {A->cx, A->cy, A->c, A->s, A->hx, A->hy}, b = {B->cx, B->cy, B->c, B->s, B->hx, B->hy};
  return box_box(&a, &b, o);
}

/* Planar inertia and floor interaction of a movable body.
 *   m     translational mass (isotropic bodies; for the ball the rolling-without-slipping effective mass 7/5 m)
 *   iz    inertia about the vertical axis
 *   flin  bound of the floor friction force (disc), ftor of the floor friction torque
 *   bfl   damping rate of the floor rows' reference acceleration (2 / (dmax * timeconst) of the floor contact)
 * The rod is anisotropic (aniso = 1): body x (across the axis) rolls -- effective mass 3/2 m, resistance
 * roll / r * N -- and body y (along the axis) slides -- mass m, bound mu N; mx/my and fx/fy hold the two axes. */
typedef struct { double m, iz, flin, ftor, bfl; int aniso; double mx, my, fx, fy; } body_par;

static void obj_mass(const orc_env* e, int type, body_par* b) {
  b->aniso = 0; b->mx = b->my = b->fx = b->fy = 0.0;
  b->bfl = 2.0 / (IMP_DMAX * SOL_TC);
  if (type == ORC_VASE || type == ORC_GREMLIN) {
    double s = type == ORC_VASE ? e->cfg.vases_size : e->cfg.gremlins_size;
    b->m = 8.0 * s * s * s * VASES_DENSITY;            /* consts.py:21,31 */
    b->iz = b->m * (2.0 / 3.0) * s * s;
    b->flin = FRICTION_MU * b->m * GRAV;
    b->ftor = b->flin * (s * sqrt(2.0));
  } else if (type == ORC_BALL) { /* dribble_ball.py:24-38 */
    double m0 = (4.0 / 3.0) * PI_D * BALL_R * BALL_R * BALL_R * BALL_DENSITY;
    b->m = 1.4 * m0;
    b->iz = 0.4 * m0 * BALL_R * BALL_R;
    b->flin = BALL_ROLL * m0 * GRAV / BALL_R;
    b->ftor = BALL_SPIN * m0 * GRAV;
    b->bfl = 2.0 / (IMP_DMAX * BALL_SOL_TC);
  } else if (type == ORC_ROD) { /* roll_rod.py:22-40 */
    double m0 = PI_D * ROD_R * ROD_R * (2.0 * ROD_HALF) * ROD_DENSITY;
    b->aniso = 1;
    b->m = m0; b->mx = 1.5 * m0; b->my = m0;
    b->iz = m0 * (3.0 * ROD_R * ROD_R + 4.0 * ROD_HALF * ROD_HALF) / 12.0;
    b->fx = ROD_ROLL * m0 * GRAV / ROD_R; b->fy = PRIO_MU * m0 * GRAV;
    b->flin = b->fy;
    b->ftor = PRIO_MU * m0 * GRAV * ROD_HALF;   /* sliding friction at the two end contacts of the line [EXT] */
  } else { /* ORC_BOX: main cube + 4 columns, push_box.py:36-67 */
    double d = BOX_SIZE, wd = BOX_SIZE / 2;
    double m0 = 8.0 * d * d * d * BOX_DENSITY, mc = 8.0 * wd * wd * d * BOX_DENSITY;
    b->m = m0 + 4.0 * mc;
    b->iz = m0 * (2.0 / 3.0) * d * d + 4.0 * (mc * (2.0 / 3.0) * wd * wd + mc * (2.0 * d * d));
    b->flin = FRICTION_MU * b->m * GRAV;
    b->ftor = b->flin * (1.5 * d);
  }
}

/* ------------------------------------------------------------------------------------------
 * Contact detection (stand-in for mj_collision [EXT]).  Canonical order: robot-vs-object in
 * slot order (robot sphere then arrow, object parts ascending), then object pairs (j asc, i<j).
 * ---------------------------------------------------------------------------------------- */
static void add_contacts(orc_env* e, int n, const double* o, int ba, int bb, int sa, int sb, int ga, int gb) {
  for (int k = 0; k < n; ++k) {
    if (e->ncon >= ORC_MAX_CON) { e->error = 1; e->overflow = 1; return; }
    orc_contact* c = &e->con[e->ncon++];
    c->ba = ba; c->bb = bb; c->sa = sa; c->sb = sb; c->ga = ga; c->gb = gb;
    c->nx = o[5 * k]; c->ny = o[5 * k + 1]; c->px = o[5 * k + 2]; c->py = o[5 * k + 3]; c->dist = o[5 * k + 4];
  }
}

static void detect(orc_env* e) {
  e->ncon = 0;
  e->overflow = 0;
  int active[ORC_MAX_OBJ];
  double o[10];
  for (int s = 0; s < e->nobj; ++s) {
    const orc_obj* ob = &e->obj[s];
    active[s] = obj_movable(ob->type) && (ob->vx != 0.0 || ob->vy != 0.0 || ob->w != 0.0);
    if (!obj_collidable(ob->type)) continue;
    for (int rg = 0; rg < robot_nparts(e); ++rg) {
      geom2 gr; robot_geom(e, rg, &gr);
      for (int p = 0; p < obj_nparts(ob->type); ++p) {
        geom2 go; obj_geom(e, s, p, &go);
        int n = collide(&gr, &go, o);
        if (n) {
          add_contacts(e, n, o, 0, obj_movable(ob->type) ? 1 + s : -1, -1, s, rg, p);
          if (obj_movable(ob->type)) active[s] = 1;
        }
      }
    }
  }
  for (int j = 0; j < e->nobj; ++j) {
    if (!obj_collidable(e->obj[j].type)) continue;
    for (int i = 0; i < j; ++i) {
      if (!obj_collidable(e->obj[i].type)) continue;
      if (!(active[i] || active[j])) continue; /* at least one awake / robot-touched movable body */
      for (int pi = 0; pi < obj_nparts(e->obj[i].type); ++pi) {
        geom2 gi; obj_geom(e, i, pi, &gi);
        for (int pj = 0; pj < obj_nparts(e->obj[j].type); ++pj) {
          geom2 gj; obj_geom(e, j, pj, &gj);
          int n = collide(&gi, &gj, o);
          if (n) add_contacts(e, n, o, obj_movable(e->obj[i].type) ? 1 + i : -1, obj_movable(e->obj[j].type) ? 1 + j : -1,
                              i, j, pi, pj);
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Point robot smooth dynamics (point.xml:15-19,35-38; SURVEY Appendix A.1/B.2-3 [EXT]).
 * Generalised coordinates are the world-frame (x, y, yaw): the two slide joints are isotropic, so
 * the joint frame rotated by robot_rot (mujoco_bridge.py:60-63) is equivalent.
 * ---------------------------------------------------------------------------------------- */
typedef struct { double p, q, ia, is; } pt_mat; /* M = [[a,0,p],[0,a,q],[p,q,I]]: ia = 1/a, is = 1/(I - (p^2+q^2)/a) */

/* the two slide joints always share one damping value, so a == b and p^2 + q^2 == (m c)^2: the Schur
 * complement is a constant of the model and the solve needs no division per substep */
static void pt_matrix(const orc_env* e, double hd, pt_mat* M) {
  if (e->robot == ORC_CAR) { /* free joint: no damping; COM offset (cx, cy) in the body frame */
    car_model c = car_params();
    double sn = sag_sin(e->q[2]), cs = sag_cos(e->q[2]);
    M->ia = 1.0 / c.M;
    M->p = -(c.mcx * sn + c.mcy * cs); M->q = c.mcx * cs - c.mcy * sn;
    M->is = 1.0 / (c.Io - (c.mcx * c.mcx + c.mcy * c.mcy) * M->ia);
    return;
  }
  double m = pt_mass(), mc = pt_mc();
  M->ia = 1.0 / (m + hd * e->damp_x);
  M->p = -mc * sag_sin(e->q[2]); M->q = mc * sag_cos(e->q[2]);
  double I = pt_inertia_o() + hd * e->damp_z;
  M->is = 1.0 / (I - mc * mc * M->ia);
}
static void pt_solve(const pt_mat* M, const double* f, double* out) {
  double al = (f[2] - (M->p * f[0] + M->q * f[1]) * M->ia) * M->is;
  out[0] = (f[0] - M->p * al) * M->ia;
  out[1] = (f[1] - M->q * al) * M->ia;
  out[2] = al;
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void pt_smooth(const orc_env* e, double* f) {
  if (e->robot == ORC_CAR) { /* chassis: no actuator, no damping; centripetal bias of the COM offset */
    car_model cm = car_params();
    double c = sag_cos(e->q[2]), s = sag_sin(e->q[2]), w = e->v[2];
    f[0] = w * w * (cm.mcx * c - cm.mcy * s);
    f[1] = w * w * (cm.mcx * s + cm.mcy * c);
    f[2] = 0.0;
    return;
  }
  double mc = pt_mc(), c = sag_cos(e->q[2]), s = sag_sin(e->q[2]), w = e->v[2];
  /* actuators: motor 'x' (site transmission, gear 0.3 along body x) and velocity servo 'z' */
  double u0 = clampd(e->ctrl[0], e->ctrl_lo[0], e->ctrl_hi[0]);
  double u1 = clampd(e->ctrl[1], e->ctrl_lo[1], e->ctrl_hi[1]);
  double fx = clampd(u0, -PT_FORCE_LIM, PT_FORCE_LIM);
  double fz = clampd(u1 - e->gear_z * w, -PT_FORCE_LIM, PT_FORCE_LIM);
  f[0] = e->gear_x * fx * c - e->damp_x * e->v[0] + mc * w * w * c;
  f[1] = e->gear_x * fx * s - e->damp_y * e->v[1] + mc * w * w * s;
  f[2] = e->gear_z * fz - e->damp_z * w;
}
/* car wheel hinges: motor gear 1, forcerange +-0.02 (car.xml:7,51-54), joint damping 0.001 (car.xml:6) */
static double car_wheel_smooth(const orc_env* e, int i) {
  double u = clampd(e->ctrl[i], e->ctrl_lo[i], e->ctrl_hi[i]);
  return clampd(u, -CAR_FORCE_LIM, CAR_FORCE_LIM) - CAR_WHEEL_DAMP * e->wheel_w[i];
}

/* ------------------------------------------------------------------------------------------
 * Soft-constraint solve (restating MuJoCo's constraint model [EXT]: reference acceleration
 * aref = -b v - k r, impedance d(r), regulariser R = (1-d)/d * A_ii; solved with a fixed number
 * of projected Gauss-Seidel sweeps).
 * ---------------------------------------------------------------------------------------- */
static double impedance(double r) {
  double x = fabs(r) / IMP_WIDTH;
  if (x > 1.0) x = 1.0;
  double y = x < 0.5 ? 2.0 * x * x : 1.0 - 2.0 * (1.0 - x) * (1.0 - x);
  return IMP_D0 + y * (IMP_DMAX - IMP_D0);
}

typedef struct {
  pt_mat M;              /* robot mass matrix (no damping) */
  double acc[1 + ORC_MAX_OBJ + 2][3]; /* robot, objects, car wheels (1 DoF each, in [.][0]) */
  double im[ORC_MAX_OBJ], ii[ORC_MAX_OBJ];
  int aniso[ORC_MAX_OBJ];            /* rod: inverse mass R diag(imx, imy) R^T = [[ma, mb], [mb, mc]] */
  double ma[ORC_MAX_OBJ], mb[ORC_MAX_OBJ], mc[ORC_MAX_OBJ];
  double iw;                         /* 1 / wheel spin inertia */
} solve_ctx;
#define WHEEL_BODY(i) (1 + ORC_MAX_OBJ + (i))

static void minv_mul(const solve_ctx* S, int body, const double* j, double* out) {
  if (body == 0) pt_solve(&S->M, j, out);
  else if (body > ORC_MAX_OBJ) { out[0] = j[0] * S->iw; out[1] = 0.0; out[2] = 0.0; }
  else if (S->aniso[body - 1]) {
    int b = body - 1;
    out[0] = S->ma[b] * j[0] + S->mb[b] * j[1]; out[1] = S->mb[b] * j[0] + S->mc[b] * j[1]; out[2] = j[2] * S->ii[b];
  }
  else { out[0] = j[0] * S->im[body - 1]; out[1] = j[1] * S->im[body - 1]; out[2] = j[2] * S->ii[body - 1]; }
}
static void body_vel(const orc_env* e, int body, double* v) {
  if (body == 0) { v[0] = e->v[0]; v[1] = e->v[1]; v[2] = e->v[2]; }
  else if (body > ORC_MAX_OBJ) { v[0] = e->wheel_w[body - 1 - ORC_MAX_OBJ]; v[1] = 0.0; v[2] = 0.0; }
  else { const orc_obj* o = &e->obj[body - 1]; v[0] = o->vx; v[1] = o->vy; v[2] = o->w; }
}
static void body_pos(const orc_env* e, int body, double* p) {
  if (body == 0) { p[0] = e->q[0]; p[1] = e->q[1]; }
  else { p[0] = e->obj[body - 1].x; p[1] = e->obj[body - 1].y; }
}

typedef struct {
  int ba, bb;
  double ja[2][3], jb[2][3]; /* row 0 normal, row 1 tangent */
  double aref[2], diag[2], R[2], f[2], inv[2];
  int type;     /* 0 contact (normal, tangent) / tendon limit, 1 equality (weld: bilateral, no projection),
                   2 wheel-floor friction (longitudinal, lateral; disc bound) */
  int nk;       /* scalar rows in this entry: 2, or 1 (tendon, weld yaw) */
  double bound; /* type 2: mu * N */
} crow;

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static void apply(solve_ctx* S, int body, const double* j, double df) {
  if (body < 0) return;
  double t[3];
  minv_mul(S, body, j, t);
  S->acc[body][0] += t[0] * df; S->acc[body][1] += t[1] * df; S->acc[body][2] += t[2] * df;
}

static void forward_dynamics(orc_env* e, double* fsmooth, double* fcon_robot, int stale_mocap) {
  solve_ctx S;
  crow rows[ORC_MAX_CON + 3];
  const int row_cap = ORC_MAX_CON + 3;
  int nrow = 0;
  int touched[ORC_MAX_OBJ];
  double ffl[ORC_MAX_OBJ][3];
  detect(e);
  pt_matrix(e, 0.0, &S.M);
  pt_smooth(e, fsmooth);
  pt_solve(&S.M, fsmooth, S.acc[0]);
  for (int s = 0; s < e->nobj; ++s) {
    S.acc[1 + s][0] = S.acc[1 + s][1] = S.acc[1 + s][2] = 0.0;
    touched[s] = 0; ffl[s][0] = ffl[s][1] = ffl[s][2] = 0.0;
    S.im[s] = S.ii[s] = 0.0; S.aniso[s] = 0;
    if (obj_movable(e->obj[s].type)) {
      body_par bp; obj_mass(e, e->obj[s].type, &bp);
      S.im[s] = 1.0 / bp.m; S.ii[s] = 1.0 / bp.iz;
      if (bp.aniso) {
        double c = sag_cos(e->obj[s].yaw), sn = sag_sin(e->obj[s].yaw), ix = 1.0 / bp.mx, iy = 1.0 / bp.my;
        S.aniso[s] = 1;
        S.ma[s] = ix * c * c + iy * sn * sn; S.mb[s] = (ix - iy) * c * sn; S.mc[s] = ix * sn * sn + iy * c * c;
      }
    }
  }
  const double bdamp = 2.0 / (IMP_DMAX * SOL_TC);
  const double kbase = 1.0 / (IMP_DMAX * IMP_DMAX * SOL_TC * SOL_TC * SOL_DR * SOL_DR);
  const double rr0 = (1.0 - IMP_D0) / IMP_D0;
  /* car: wheel-floor friction, always present.  Reduced model: the chassis stays level, normal loads
   * are the static ones, each wheel has a longitudinal slip row (chassis point velocity along the rolling direction +
   * r * wheel rate) and a lateral slip row, jointly bounded by mu * N; the rear castor rolls ideally. */
  if (e->robot == ORC_CAR) {
    car_model cm = car_params();
    S.iw = 1.0 / cm.Iw;
    double cs = sag_cos(e->q[2]), sn = sag_sin(e->q[2]);
    for (int i = 0; i < 2; ++i) {
      S.acc[WHEEL_BODY(i)][0] = car_wheel_smooth(e, i) * S.iw; S.acc[WHEEL_BODY(i)][1] = S.acc[WHEEL_BODY(i)][2] = 0.0;
      double bx = i == 0 ? -0.1 : 0.1, by = 0.1;              /* wheel bodies, car.xml:21,25 */
      double rx = bx * cs - by * sn, ry = bx * sn + by * cs;  /* lever arm in the world frame */
      crow* r = &rows[nrow++];
      r->type = 2; r->nk = 2; r->bound = FRICTION_MU * cm.nwheel;
      r->ba = 0; r->bb = WHEEL_BODY(i);
      r->ja[0][0] = -sn; r->ja[0][1] = cs; r->ja[0][2] = rx * cs - ry * -sn;  /* rolling direction = body y */
      r->jb[0][0] = CAR_WHEEL_R; r->jb[0][1] = 0.0; r->jb[0][2] = 0.0;
      r->ja[1][0] = cs; r->ja[1][1] = sn; r->ja[1][2] = rx * sn - ry * cs;    /* lateral = body x */
      r->jb[1][0] = 0.0; r->jb[1][1] = 0.0; r->jb[1][2] = 0.0;
      double va[3], vb[3];
      body_vel(e, 0, va); body_vel(e, r->bb, vb);
      for (int k = 0; k < 2; ++k) {
        double t[3], diag = 0.0, vel = 0.0;
        minv_mul(&S, 0, r->ja[k], t); diag += dot3(r->ja[k], t); vel += dot3(r->ja[k], va);
        if (k == 0) { minv_mul(&S, r->bb, r->jb[k], t); diag += dot3(r->jb[k], t); vel += dot3(r->jb[k], vb); }
        r->diag[k] = diag; r->R[k] = rr0 * diag; r->aref[k] = -bdamp * vel; r->f[k] = 0.0;
      }
    }
  }
  /* contact rows (active only when penetrating: dist < 0) */
  for (int i = 0; i < e->ncon; ++i) {
    const orc_contact* c = &e->con[i];
    if (!(c->dist < 0.0)) continue;
    if (c->ba < 0 && c->bb < 0) continue;
    if (nrow >= row_cap) { e->error = 1; e->overflow = 1; break; }  /* only reachable with weld rows in the table */
    crow* r = &rows[nrow++];
    r->type = 0; r->nk = 2; r->bound = FRICTION_MU;
    r->ba = c->ba; r->bb = c->bb;
    /* a ball / rod geom has priority 1 (dribble_ball.py:38, roll_rod.py:40): the contact takes its friction and
     * solref instead of the max / default mix [EXT] */
    double cb = bdamp, ck = kbase;
    {
      int ta = c->sa >= 0 ? e->obj[c->sa].type : ORC_NONE, tb = c->sb >= 0 ? e->obj[c->sb].type : ORC_NONE;
      if (ta == ORC_ROD || tb == ORC_ROD || ta == ORC_BALL || tb == ORC_BALL) r->bound = PRIO_MU;
      if (ta == ORC_BALL || tb == ORC_BALL) {
        cb = 2.0 / (IMP_DMAX * BALL_SOL_TC);
        ck = 1.0 / (IMP_DMAX * IMP_DMAX * BALL_SOL_TC * BALL_SOL_TC * BALL_SOL_DR * BALL_SOL_DR);
      }
    }
    double tx = -c->ny, ty = c->nx;
    double pa[2] = {0, 0}, pb[2] = {0, 0}, va[3] = {0, 0, 0}, vb[3] = {0, 0, 0};
    if (c->ba >= 0) { body_pos(e, c->ba, pa); body_vel(e, c->ba, va); if (c->ba > 0) touched[c->ba - 1] = 1; }
    if (c->bb >= 0) { body_pos(e, c->bb, pb); body_vel(e, c->bb, vb); if (c->bb > 0) touched[c->bb - 1] = 1; }
    double rax = c->px - pa[0], ray = c->py - pa[1], rbx = c->px - pb[0], rby = c->py - pb[1];
    r->ja[0][0] = -c->nx; r->ja[0][1] = -c->ny; r->ja[0][2] = -(rax * c->ny - ray * c->nx);
    r->jb[0][0] = c->nx; r->jb[0][1] = c->ny; r->jb[0][2] = rbx * c->ny - rby * c->nx;
    r->ja[1][0] = -tx; r->ja[1][1] = -ty; r->ja[1][2] = -(rax * ty - ray * tx);
    r->jb[1][0] = tx; r->jb[1][1] = ty; r->jb[1][2] = rbx * ty - rby * tx;
    double d = impedance(c->dist);
    for (int k = 0; k < 2; ++k) {
      double t[3], diag = 0.0, vel = 0.0;
      if (c->ba >= 0) { minv_mul(&S, c->ba, r->ja[k], t); diag += dot3(r->ja[k], t); vel += dot3(r->ja[k], va); }
      if (c->bb >= 0) { minv_mul(&S, c->bb, r->jb[k], t); diag += dot3(r->jb[k], t); vel += dot3(r->jb[k], vb); }
      r->diag[k] = diag;
      r->R[k] = (1.0 - d) / d * diag;
      r->aref[k] = -cb * vel - (k == 0 ? d * ck * c->dist : 0.0);
      r->f[k] = 0.0;
    }
  }
  /* HaulBox tendon length limit (haul_box.py:21-30): site 'robot' <-> 'box_site', range [0, 0.75] */
  int tendon_row = -1;
  if (e->tendon_slot >= 0) {
    const orc_obj* b = &e->obj[e->tendon_slot];
    double dx = b->x - e->q[0], dy = b->y - e->q[1], dz = BOX_SIZE - PT_Z;
    double len = sqrt(dx * dx + dy * dy + dz * dz);
    double dist = TENDON_MAX - len;
    if (dist < 0.0) {
      if (nrow >= row_cap) { e->error = 1; e->overflow = 1; }
      else {
      crow* r = &rows[nrow]; tendon_row = nrow++;
      r->type = 0; r->nk = 1; r->bound = 0.0;
      r->ba = 0; r->bb = 1 + e->tendon_slot;
      /* d(dist)/dq : robot +d/len, box -d/len */
      r->ja[0][0] = dx / len; r->ja[0][1] = dy / len; r->ja[0][2] = 0.0;
      r->jb[0][0] = -dx / len; r->jb[0][1] = -dy / len; r->jb[0][2] = 0.0;
      double va[3], vb[3], t[3], diag = 0.0;
      body_vel(e, 0, va); body_vel(e, r->bb, vb);
      minv_mul(&S, 0, r->ja[0], t); diag += dot3(r->ja[0], t);
      minv_mul(&S, r->bb, r->jb[0], t); diag += dot3(r->jb[0], t);
      double d = impedance(dist);
      r->diag[0] = diag; r->R[0] = (1.0 - d) / d * diag;
      r->aref[0] = -bdamp * (dot3(r->ja[0], va) + dot3(r->jb[0], vb)) - d * kbase * dist;
      r->f[0] = 0.0;
      touched[e->tendon_slot] = 1;
      }
    }
  }
  /* gremlin welds (primitive_objects.py:79-82, world.py:157-165): a soft equality between the gremlin and its mocap
   * body, target = spawn pose + mocap position [EXT], solref (0.02, 1.5), default solimp; planar reduction: one
   * bilateral row pair (x, y) and one bilateral row (yaw) per gremlin, impedance per row from its own residual.  The
   * first substep of physics.step still sees the mocap position of the previous kinematics pass. */
  {
    const double* mp = stale_mocap ? e->mocap_kin : e->mocap_pos;
    const double wb_ = 2.0 / (IMP_DMAX * WELD_SOL_TC);
    const double wk_ = 1.0 / (IMP_DMAX * IMP_DMAX * WELD_SOL_TC * WELD_SOL_TC * WELD_SOL_DR * WELD_SOL_DR);
    for (int s = 0; s < e->nobj; ++s) {
      if (e->obj[s].type != ORC_GREMLIN) continue;
      if (nrow + 2 > row_cap) { e->error = 1; e->overflow = 1; break; }
      const orc_obj* o = &e->obj[s];
      double res[3] = {o->x - (e->gspawn[s][0] + mp[0]), o->y - (e->gspawn[s][1] + mp[1]), o->yaw - e->gspawn[s][2]};
      double vb[3]; body_vel(e, 1 + s, vb);
      for (int part = 0; part < 2; ++part) {
        crow* r = &rows[nrow++];
        r->type = 1; r->nk = part == 0 ? 2 : 1; r->bound = 0.0;
        r->ba = -1; r->bb = 1 + s;
        for (int k = 0; k < 2; ++k) for (int d = 0; d < 3; ++d) { r->ja[k][d] = 0.0; r->jb[k][d] = 0.0; }
        if (part == 0) { r->jb[0][0] = 1.0; r->jb[1][1] = 1.0; } else r->jb[0][2] = 1.0;
        for (int k = 0; k < r->nk; ++k) {
          double t[3], resid = part == 0 ? res[k] : res[2];
          minv_mul(&S, r->bb, r->jb[k], t);
          double diag = dot3(r->jb[k], t), d = impedance(resid);
          r->diag[k] = diag; r->R[k] = (1.0 - d) / d * diag;
          r->aref[k] = -wb_ * dot3(r->jb[k], vb) - d * wk_ * resid;
          r->f[k] = 0.0;
        }
      }
      touched[s] = 1;
    }
    if (!stale_mocap) { e->mocap_kin[0] = e->mocap_pos[0]; e->mocap_kin[1] = e->mocap_pos[1]; }
  }
  /* floor-friction rows for awake / touched movable bodies (plane contact reduced to the plane) */
  int nfl = 0, flslot[ORC_MAX_OBJ];
  for (int s = 0; s < e->nobj; ++s) {
    const orc_obj* o = &e->obj[s];
    if (!obj_movable(o->type)) continue;
    if (touched[s] || o->vx != 0.0 || o->vy != 0.0 || o->w != 0.0) flslot[nfl++] = s;
  }
  /* capacity limits (the GPU keeps the solver's working set in shared memory): on overflow the step is a
   * PhysicsError and this pass applies no constraint forces and moves no object */
  if (nfl > ORC_MAX_BODIES) { e->error = 1; e->overflow = 1; }
  if (e->overflow) { nrow = e->robot == ORC_CAR ? 2 : 0; nfl = 0; tendon_row = -1; for (int s = 0; s < e->nobj; ++s) touched[s] = 0; }
  const double rr = (1.0 - IMP_D0) / IMP_D0;
  /* Projected Gauss-Seidel.  Row update: f <- proj(f - (J a - aref + R f) / (A_ii + R)); the reciprocal of the
   * denominator is taken once per row.  Terminates after PGS_SWEEPS sweeps or when one sweep changes the forces by
   * less than PGS_TOL relative (L1) -- MuJoCo's own solvers stop at `tolerance` = 1e-8 [EXT]. */
  for (int i = 0; i < nrow; ++i) {
    for (int k = 0; k < rows[i].nk; ++k) rows[i].inv[k] = 1.0 / (rows[i].diag[k] + rows[i].R[k]);
  }
  (void)tendon_row;
  for (int it = 0; it < PGS_SWEEPS; ++it) {
    double sdf = 0.0, sf = 0.0;
    for (int i = 0; i < nrow; ++i) {
      crow* r = &rows[i];
      int nk = r->nk;
      if (r->type == 2) { /* wheel: both slip rows, then project the pair onto the friction disc */
        double fo[2] = {r->f[0], r->f[1]};
        for (int k = 0; k < 2; ++k) {
          double a = dot3(r->ja[k], S.acc[r->ba]);
          if (k == 0) a += dot3(r->jb[k], S.acc[r->bb]);
          double fn = r->f[k] - (a - r->aref[k] + r->R[k] * r->f[k]) * r->inv[k];
          double df = fn - r->f[k];
          r->f[k] = fn;
          if (df != 0.0) { apply(&S, r->ba, r->ja[k], df); if (k == 0) apply(&S, r->bb, r->jb[k], df); }
        }
        double nf = sqrt(r->f[0] * r->f[0] + r->f[1] * r->f[1]);
        if (nf > r->bound) {
          double sc = r->bound / nf;
          for (int k = 0; k < 2; ++k) {
            double g = r->f[k] * sc, df = g - r->f[k];
            r->f[k] = g;
            if (df != 0.0) { apply(&S, r->ba, r->ja[k], df); if (k == 0) apply(&S, r->bb, r->jb[k], df); }
          }
        }
        sdf += fabs(r->f[0] - fo[0]) + fabs(r->f[1] - fo[1]); sf += fabs(r->f[0]) + fabs(r->f[1]);
        continue;
      }
      for (int k = 0; k < nk; ++k) {
        double a = 0.0;
        if (r->ba >= 0) a += dot3(r->ja[k], S.acc[r->ba]);
        if (r->bb >= 0) a += dot3(r->jb[k], S.acc[r->bb]);
        double fn = r->f[k] - (a - r->aref[k] + r->R[k] * r->f[k]) * r->inv[k];
        if (r->type == 1) { /* equality: bilateral, unbounded */ }
        else if (k == 0) { if (fn < 0.0) fn = 0.0; }
        else { double lim = r->bound * r->f[0]; fn = clampd(fn, -lim, lim); }
        double df = fn - r->f[k];
        r->f[k] = fn;
        sdf += fabs(df); sf += fabs(fn);
        if (df != 0.0) { apply(&S, r->ba, r->ja[k], df); apply(&S, r->bb, r->jb[k], df); }
      }
    }
    for (int i = 0; i < nfl; ++i) {
      int s = flslot[i];
      const orc_obj* o = &e->obj[s];
      body_par bp; obj_mass(e, o->type, &bp);
      double* ac = S.acc[1 + s];
      double At = S.ii[s];
      double inv_tor = 1.0 / (At + rr * At);
      double f0, f1, d0, d1;
      if (bp.aniso) {
        /* rod: one row across the axis (rolling, body x = u) and one along it (sliding, body y = w), each an
         * eigen-direction of the inverse mass, clamped separately */
        double c = sag_cos(o->yaw), sn = sag_sin(o->yaw);
        double ix = 1.0 / bp.mx, iy = 1.0 / bp.my;
        double inv_x = 1.0 / (ix + rr * ix), inv_y = 1.0 / (iy + rr * iy);
        double au = ac[0] * c + ac[1] * sn, aw = -ac[0] * sn + ac[1] * c;
        double vu = o->vx * c + o->vy * sn, vw = -o->vx * sn + o->vy * c;
        f0 = ffl[s][0] - (au + bp.bfl * vu + rr * ix * ffl[s][0]) * inv_x;
        f1 = ffl[s][1] - (aw + bp.bfl * vw + rr * iy * ffl[s][1]) * inv_y;
        f0 = clampd(f0, -bp.fx, bp.fx); f1 = clampd(f1, -bp.fy, bp.fy);
        d0 = f0 - ffl[s][0]; d1 = f1 - ffl[s][1];
        double du = d0 * ix, dw = d1 * iy;
        ac[0] += du * c - dw * sn; ac[1] += du * sn + dw * c;
      } else {
        double Al = S.im[s];
        double inv_lin = 1.0 / (Al + rr * Al);
        f0 = ffl[s][0] - (ac[0] + bp.bfl * o->vx + rr * Al * ffl[s][0]) * inv_lin;
        f1 = ffl[s][1] - (ac[1] + bp.bfl * o->vy + rr * Al * ffl[s][1]) * inv_lin;
        double nf = sqrt(f0 * f0 + f1 * f1);
        if (nf > bp.flin) { double sc = bp.flin / nf; f0 *= sc; f1 *= sc; }
        d0 = f0 - ffl[s][0]; d1 = f1 - ffl[s][1];
        ac[0] += d0 * Al; ac[1] += d1 * Al;
      }
      ffl[s][0] = f0; ffl[s][1] = f1;
      double f2 = ffl[s][2] - (ac[2] + bp.bfl * o->w + rr * At * ffl[s][2]) * inv_tor;
      f2 = clampd(f2, -bp.ftor, bp.ftor);
      double d2 = f2 - ffl[s][2];
      ac[2] += d2 * At;
      ffl[s][2] = f2;
      sdf += fabs(d0) + fabs(d1) + fabs(d2); sf += fabs(f0) + fabs(f1) + fabs(f2);
    }
    if (sdf <= PGS_TOL * sf) break;
  }
  /* results */
  for (int k = 0; k < 3; ++k) e->qacc[k] = S.acc[0][k];
  fcon_robot[0] = fcon_robot[1] = fcon_robot[2] = 0.0;
  for (int i = 0; i < nrow; ++i) {
    crow* r = &rows[i];
    int nk = r->nk;
    for (int k = 0; k < nk; ++k) {
      if (r->ba == 0) for (int d = 0; d < 3; ++d) fcon_robot[d] += r->ja[k][d] * r->f[k];
      if (r->bb == 0) for (int d = 0; d < 3; ++d) fcon_robot[d] += r->jb[k][d] * r->f[k];
    }
  }
  e->wheel_tau[0] = e->wheel_tau[1] = 0.0;
  for (int i = 0; i < nrow; ++i) if (rows[i].type == 2) e->wheel_tau[rows[i].bb - WHEEL_BODY(0)] = rows[i].jb[0][0] * rows[i].f[0];
  for (int s = 0; s < e->nobj; ++s) {
    e->oacc[s][0] = S.acc[1 + s][0]; e->oacc[s][1] = S.acc[1 + s][1]; e->oacc[s][2] = S.acc[1 + s][2];
    e->touched[s] = touched[s];
  }
}

void orc_phys_forward(orc_env* e) {
  double fs[3], fc[3];
  forward_dynamics(e, fs, fc, 0);
}

static int bad(double x) { return !(fabs(x) <= 1e10); } /* NaN or > mjMAXVAL [EXT] */

/* castor ball (car.xml:29-32): ideal rolling.  Relative angular velocity of the ball joint in the child frame, from
 * the chassis point velocity at the castor (mjSENS_BALLANGVEL semantics [EXT]). */
static void car_castor_angvel(const orc_env* e, double* wc) {
  double cs = sag_cos(e->q[2]), sn = sag_sin(e->q[2]);
  double bx = 0.0, by = -0.1;
  double rx = bx * cs - by * sn, ry = bx * sn + by * cs;
  double vx = e->v[0] - e->v[2] * ry, vy = e->v[1] + e->v[2] * rx;  /* world velocity of the castor centre */
  double wx = -vy / CAR_WHEEL_R, wy = vx / CAR_WHEEL_R;              /* rolling without slip; no spin relative to the chassis */
  double px = wx * cs + wy * sn, py = -wx * sn + wy * cs, pz = 0.0;  /* parent (chassis) frame */
  /* child frame: rotate by the inverse of the joint quaternion */
  double w = e->cq[0], x = -e->cq[1], y = -e->cq[2], z = -e->cq[3];
  double tx = 2.0 * (y * pz - z * py), ty = 2.0 * (z * px - x * pz), tz = 2.0 * (x * py - y * px);
  wc[0] = px + w * tx + (y * tz - z * ty);
  wc[1] = py + w * ty + (z * tx - x * tz);
  wc[2] = pz + w * tz + (x * ty - y * tx);
}
/* wheels: implicit joint damping; castor: quaternion exponential-map update (mju_quatIntegrate [EXT]) */
static void car_integrate_extra(orc_env* e) {
  car_model cm = car_params();
  for (int i = 0; i < 2; ++i) {
    double al = (car_wheel_smooth(e, i) + e->wheel_tau[i]) / (cm.Iw + e->h * CAR_WHEEL_DAMP);
    e->wheel_w[i] += e->h * al;
    if (bad(e->wheel_w[i])) e->error = 1;
  }
  double wc[3];
  car_castor_angvel(e, wc);
  double wn = sqrt(wc[0] * wc[0] + wc[1] * wc[1] + wc[2] * wc[2]);
  if (wn > 0.0) {
    double sh, ch;
    sag_sincos(0.5 * e->h * wn, &sh, &ch);
    double ax = wc[0] / wn * sh, ay = wc[1] / wn * sh, az = wc[2] / wn * sh;
    double a = e->cq[0], b = e->cq[1], c = e->cq[2], d = e->cq[3];
    double q0 = a * ch - b * ax - c * ay - d * az;
    double q1 = a * ax + b * ch + c * az - d * ay;
    double q2 = a * ay - b * az + c * ch + d * ax;
    double q3 = a * az + b * ay - c * ax + d * ch;
    double n = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    e->cq[0] = q0 / n; e->cq[1] = q1 / n; e->cq[2] = q2 / n; e->cq[3] = q3 / n;
  }
}

/* physics.step(nstep): nstep x (forward dynamics, semi-implicit Euler with implicit joint
 * damping), safe_adaptation_gym.py:72; SURVEY Appendix B.1-2 [EXT] */
void orc_phys_step(orc_env* e, int nstep) {
  for (int it = 0; it < nstep; ++it) {
    double fs[3], fc[3], rhs[3], a[3];
    forward_dynamics(e, fs, fc, it == 0);  /* mj_step2 first: the mocap edit is not seen yet (App. B.1) */
    pt_mat Mh; pt_matrix(e, e->h, &Mh);
    for (int k = 0; k < 3; ++k) rhs[k] = fs[k] + fc[k];
    pt_solve(&Mh, rhs, a);
    for (int k = 0; k < 3; ++k) { e->v[k] += e->h * a[k]; }
    for (int k = 0; k < 3; ++k) { e->q[k] += e->h * e->v[k]; }
    if (e->robot == ORC_CAR) car_integrate_extra(e);
    for (int s = 0; s < e->nobj; ++s) {
      orc_obj* o = &e->obj[s];
      if (!obj_movable(o->type)) continue;
      int touched = e->touched[s];
      if (!(touched || o->vx != 0.0 || o->vy != 0.0 || o->w != 0.0)) continue;
      if (e->overflow) continue;
      o->vx += e->h * e->oacc[s][0]; o->vy += e->h * e->oacc[s][1]; o->w += e->h * e->oacc[s][2];
      if (!touched && o->vx * o->vx + o->vy * o->vy < SLEEP_V * SLEEP_V && fabs(o->w) < SLEEP_V) { o->vx = o->vy = o->w = 0.0; }
      o->x += e->h * o->vx; o->y += e->h * o->vy; o->yaw += e->h * o->w;
      if (bad(o->x) || bad(o->y) || bad(o->vx) || bad(o->vy) || bad(o->w)) e->error = 1;
    }
    e->time += e->h;
    for (int k = 0; k < 3; ++k) if (bad(e->q[k]) || bad(e->v[k]) || bad(a[k])) e->error = 1;
  }
}

void orc_phys_set_control(orc_env* e, const double* u) { e->ctrl[0] = u[0]; e->ctrl[1] = u[1]; }
int orc_phys_ncon(const orc_env* e) { return e->ncon; }
void orc_phys_get_contact(const orc_env* e, int i, orc_contact* out) { *out = e->con[i]; }
int orc_phys_error(const orc_env* e) { return e->error; }
double orc_phys_time(const orc_env* e) { return e->time; }

/* _sensors(), safe_adaptation_gym.py:225-237; sensor semantics SURVEY Appendix B.6 [EXT] */
void orc_phys_sensors(const orc_env* e, double* out) {
  double c = sag_cos(e->q[2]), s = sag_sin(e->q[2]);
  out[0] = e->qacc[0] * c + e->qacc[1] * s;      /* accelerometer */
  out[1] = -e->qacc[0] * s + e->qacc[1] * c;
  out[2] = GRAV;
  out[3] = e->v[0] * c + e->v[1] * s;            /* velocimeter */
  out[4] = -e->v[0] * s + e->v[1] * c;
  out[5] = 0.0;
  out[6] = 0.0; out[7] = 0.0; out[8] = e->v[2];  /* gyro */
  out[9] = -0.5 * s; out[10] = -0.5 * c; out[11] = 0.0; /* magnetometer: R^T (0,-0.5,0) */
  if (e->robot == ORC_CAR) { /* ballangvel_rear (3) then quat2mat(ballquat_rear).ravel() (9): safe_adaptation_gym.py:228-236 */
    car_castor_angvel(e, out + 12);
    double w = e->cq[0], x = e->cq[1], y = e->cq[2], z = e->cq[3];
    out[15] = w * w + x * x - y * y - z * z; out[16] = 2.0 * (x * y - w * z); out[17] = 2.0 * (x * z + w * y);
    out[18] = 2.0 * (x * y + w * z); out[19] = w * w - x * x + y * y - z * z; out[20] = 2.0 * (y * z - w * x);
    out[21] = 2.0 * (x * z - w * y); out[22] = 2.0 * (y * z + w * x); out[23] = w * w - x * x - y * y + z * z;
  }
}

/* ==========================================================================================
 * Lidar: SafeAdaptationGym._lidar, safe_adaptation_gym.py:174-223
 * ======================================================================================== */
/* literal restatement, line by line; pinned by tests/golden/lidar_kat.json */
static void lidar_accum_literal(double rx, double ry, double c, double s, double px, double py, double* obs) {
  /* ego_xy (:197-202): (pos - robot_pos) @ robot_mat, planar */
  double wx = px - rx, wy = py - ry;
  double ex = wx * c + wy * s, ey = -wx * s + wy * c;
  double dist = sqrt(ex * ex + ey * ey);                      /* :209 np.abs(z) */
  double angle = sag_atan2(ey, ex);                               /* :210 np.angle(z) % 2pi */
  if (angle < 0.0) angle += TWO_PI;
  const double inv_bin_size = ORC_NUM_LIDAR_BINS / TWO_PI;    /* :211 (1 / bin_size) */
  double t = angle * inv_bin_size;
  int bin = (int)t;                                           /* :212 */
  double sensor = LIDAR_MAX_DIST - dist;                      /* :214 */
  if (sensor < 0.0) sensor = 0.0;
  sensor *= 1.0 / LIDAR_MAX_DIST;
  double alias = t - (double)bin;                             /* :213,216: (angle - bin*bin_size)/bin_size */
  int b0 = bin % ORC_NUM_LIDAR_BINS;                          /* quirk D7: wrap instead of IndexError */
  int bp = (bin + 1) % ORC_NUM_LIDAR_BINS, bm = (bin + ORC_NUM_LIDAR_BINS - 1) % ORC_NUM_LIDAR_BINS;
  if (sensor > obs[b0]) obs[b0] = sensor;                     /* :215 */
  if (alias * sensor > obs[bp]) obs[bp] = alias * sensor;     /* :221 */
  if ((1.0 - alias) * sensor > obs[bm]) obs[bm] = (1.0 - alias) * sensor; /* :222 */
}

/* the form the environment (and the GPU, bit for bit) evaluates: same quantities, with (bin, alias) from
 * sag_lidar_bin16 (one division, no full atan2 -- include/sag_detmath.h) and fused multiply-adds in the ego transform.
 * Pinned by the same known answers and against the literal form on random inputs (tests/test_golden.py,
 * tests/test_detmath.py). */
static void lidar_accum(double rx, double ry, double c, double s, double px, double py, double* obs) {
  double wx = px - rx, wy = py - ry;
  double ex = SAG_FMA(wx, c, wy * s), ey = SAG_FMA(wy, c, -(wx * s)); /* :197-202 */
  double dist = sqrt(SAG_FMA(ex, ex, ey * ey));                       /* :209 */
  int bin;
  double alias;
  sag_lidar_bin16(ex, ey, &bin, &alias);                              /* :210-213,216 */
  double sensor = LIDAR_MAX_DIST - dist;                              /* :214 */
  if (sensor < 0.0) sensor = 0.0;
  sensor *= 1.0 / LIDAR_MAX_DIST;
  int b0 = bin & (ORC_NUM_LIDAR_BINS - 1);                            /* quirk D7: wrap instead of IndexError */
  int bp = (bin + 1) & (ORC_NUM_LIDAR_BINS - 1), bm = (bin + ORC_NUM_LIDAR_BINS - 1) & (ORC_NUM_LIDAR_BINS - 1);
  if (sensor > obs[b0]) obs[b0] = sensor;                             /* :215 */
  if (alias * sensor > obs[bp]) obs[bp] = alias * sensor;             /* :221 */
  if ((1.0 - alias) * sensor > obs[bm]) obs[bm] = (1.0 - alias) * sensor; /* :222 */
}

void orc_lidar(double rx, double ry, double ryaw, int n, const double* xs, const double* ys, double* out16) {
  double c = sag_cos(ryaw), s = sag_sin(ryaw);
  for (int i = 0; i < ORC_NUM_LIDAR_BINS; ++i) out16[i] = 0.0;
  for (int i = 0; i < n; ++i) lidar_accum(rx, ry, c, s, xs[i], ys[i], out16);
}
void orc_lidar_literal(double rx, double ry, double ryaw, int n, const double* xs, const double* ys, double* out16) {
  double c = sag_cos(ryaw), s = sag_sin(ryaw);
  for (int i = 0; i < ORC_NUM_LIDAR_BINS; ++i) out16[i] = 0.0;
  for (int i = 0; i < n; ++i) lidar_accum_literal(rx, ry, c, s, xs[i], ys[i], out16);
}

/* observation: safe_adaptation_gym.py:120-139 -- [obstacles(16), objects(16), goal(16), sensors] */
void orc_env_observation(orc_env* e, double* obs) {
  double c = sag_cos(e->q[2]), s = sag_sin(e->q[2]);
  for (int i = 0; i < 48; ++i) obs[i] = 0.0;
  for (int k = 0; k < e->nobj; ++k) {
    const orc_obj* o = &e->obj[k];
    double* dst = 0;
    if (o->group == ORC_GROUP_OBSTACLES) dst = obs;           /* world.py:225-226 */
    else if (o->group == ORC_GROUP_OBJECTS) dst = obs + 16;   /* world.py:229-230 */
    else if (o->group == ORC_GROUP_GOAL) dst = obs + 32;      /* world.py:227-228 */
    if (dst) lidar_accum(e->q[0], e->q[1], c, s, o->x, o->y, dst);
  }
  orc_phys_sensors(e, obs + 48);
}
int orc_env_obs_dim(const orc_env* e) { return e->robot == ORC_CAR ? 72 : 60; }

/* ==========================================================================================
 * Task table (tasks/ *.py)
 * ======================================================================================== */
typedef struct {
  int obstacles[4];     /* hazards, vases, gremlins, pillars */
  double extent;        /* placement_extents = (-e,-e,e,e) */
  int kind;             /* 0 goal-only, 1 buttons, 2 goal+box */
  int nbuttons;
  double button_rect;   /* buttons rect half-size */
  int box_type;         /* ORC_BOX / ORC_ROD / ORC_BALL */
  double box_keepout;
  double box_rect;      /* 0 = free placement */
} task_spec;

static const task_spec TASKS[ORC_NUM_TASKS] = {
    /* catch_goal          */ {{9, 10, 0, 1}, 2.0, 0, 0, 0, 0, 0, 0},          /* go_to_goal.py:84 (inherited) */
    /* collect             */ {{6, 8, 0, 0}, 2.25, 1, 6, 1.5, 0, 0, 0},        /* press_buttons.py:98 (inherited), collect.py:11,21,50-51 */
    /* dribble_ball        */ {{2, 3, 0, 1}, 1.75, 2, 0, 0, ORC_BALL, 0.2, 0}, /* dribble_ball.py:11-13,43-45 */
    /* go_to_goal          */ {{9, 10, 0, 1}, 2.0, 0, 0, 0, 0, 0, 0},          /* go_to_goal.py:84 */
    /* go_to_goal_damping  */ {{9, 10, 0, 1}, 2.0, 0, 0, 0, 0, 0, 0},
    /* go_to_goal_motor    */ {{9, 10, 0, 1}, 2.0, 0, 0, 0, 0, 0, 0},
    /* go_to_goal_scarce   */ {{9, 10, 0, 1}, 2.0, 0, 0, 0, 0, 0, 0},
    /* haul_box            */ {{2, 3, 0, 1}, 1.75, 2, 0, 0, ORC_BOX, 0.5, 0},  /* push_box.py:13,104-108 */
    /* press_buttons       */ {{6, 8, 0, 0}, 2.0, 1, 4, 1.35, 0, 0, 0},        /* press_buttons.py:13,29,98 */
    /* press_buttons_scarce*/ {{6, 8, 0, 0}, 2.0, 1, 4, 1.75, 0, 0, 0},        /* press_buttons_scarce.py:17 */
    /* push_box            */ {{2, 3, 0, 1}, 1.75, 2, 0, 0, ORC_BOX, 0.5, 0},
    /* push_box_scarce     */ {{2, 3, 0, 1}, 1.75, 2, 0, 0, ORC_BOX, 0.55, 2.25}, /* push_box_scarce.py:18 */
    /* roll_rod            */ {{2, 3, 0, 1}, 1.75, 2, 0, 0, ORC_ROD, 0.7, 0},  /* roll_rod.py:14 */
    /* unsupervised        */ {{5, 6, 0, 1}, 2.0, 0, 0, 0, 0, 0, 0},           /* unsupervised.py:80 */
};

static int task_nobj_g(int task, int ng) {
  const task_spec* t = &TASKS[task];
  int n = t->obstacles[0] + t->obstacles[1] + ng + t->obstacles[3];
  if (t->kind == 0) n += 1; else if (t->kind == 1) n += t->nbuttons; else n += 2;
  return n;
}
int orc_task_nobj(int task) { return task_nobj_g(task, TASKS[task].obstacles[2]); }
static void task_slot_types_g(int task, int ng, int* types);
void orc_task_slot_types(int task, int* types) { task_slot_types_g(task, TASKS[task].obstacles[2], types); }
/* ng: Task.obstacles[2] -- 0 in every shipped task (task.py:70 and the overrides), set through orc_config.num_gremlins */
static void task_slot_types_g(int task, int ng, int* types) {
  const task_spec* t = &TASKS[task];
  static const int kinds[4] = {ORC_HAZARD, ORC_VASE, ORC_GREMLIN, ORC_PILLAR};
  int n = 0;
  for (int k = 0; k < 4; ++k) for (int i = 0; i < (k == 2 ? ng : t->obstacles[k]); ++i) types[n++] = kinds[k];
  if (t->kind == 0) types[n++] = ORC_GOAL;
  else if (t->kind == 1) for (int i = 0; i < t->nbuttons; ++i) types[n++] = ORC_BUTTON;
  else { types[n++] = ORC_GOAL; types[n++] = t->box_type; }
  for (; n < ORC_MAX_OBJ; ++n) types[n] = ORC_NONE;
}

void orc_default_config(orc_config* c) { /* world.py:17-34 */
  c->placements_margin = 0.0; c->robot_keepout = 0.4;
  c->hazards_size = 0.2; c->vases_size = 0.1; c->pillars_size = 0.2; c->gremlins_size = 0.1;
  c->hazards_keepout = 0.18; c->gremlins_keepout = 0.4; c->vases_keepout = 0.15; c->pillars_keepout = 0.3;
  c->gremlins_travel = 0.35; c->robot_ctrl_range_scale = 0.0; c->action_noise = 0.01; c->max_bound = 25.0;
  c->random_bound = 0; c->max_layout_draws = 0; c->num_gremlins = 0;
}

static void setup_slots(orc_env* e) {
  /* world.py:53-102 (_setup_placements, keepouts) + task.setup_placements */
  const task_spec* t = &TASKS[e->task];
  int types[ORC_MAX_OBJ];
  task_slot_types_g(e->task, e->cfg.num_gremlins, types);
  e->nobj = task_nobj_g(e->task, e->cfg.num_gremlins);
  e->goal_slot = e->box_slot = e->first_button = -1;
  e->nbuttons = t->nbuttons;
  e->tendon_slot = -1;
  double ext = t->extent;
  e->extents[0] = -ext; e->extents[1] = -ext; e->extents[2] = ext; e->extents[3] = ext;
  const orc_config* c = &e->cfg;
  for (int s = 0; s < e->nobj; ++s) {
    orc_obj* o = &e->obj[s];
    memset(o, 0, sizeof(*o));
    o->type = types[s];
    e->has_rect[s] = 0;
    switch (o->type) {
      case ORC_HAZARD: o->keepout = c->hazards_keepout < c->hazards_size ? c->hazards_size : c->hazards_keepout; o->group = ORC_GROUP_OBSTACLES; break;
      case ORC_VASE: o->keepout = c->vases_keepout < c->vases_size ? c->vases_size : c->vases_keepout; o->group = ORC_GROUP_OBSTACLES; break;
      case ORC_GREMLIN: o->keepout = c->gremlins_keepout < c->gremlins_size ? c->gremlins_size : c->gremlins_keepout; o->group = ORC_GROUP_OBSTACLES; break;
      case ORC_PILLAR: o->keepout = c->pillars_keepout < c->pillars_size ? c->pillars_size : c->pillars_keepout; o->group = ORC_GROUP_OBSTACLES; break;
      case ORC_GOAL:
        o->keepout = GOAL_KEEPOUT; o->group = ORC_GROUP_GOAL; e->goal_slot = s;
        e->has_rect[s] = 1; e->rect[s][0] = e->rect[s][1] = -1.5; e->rect[s][2] = e->rect[s][3] = 1.5; /* go_to_goal.py:9 */
        break;
      case ORC_BUTTON:
        o->keepout = BUTTONS_KEEPOUT; o->group = ORC_GROUP_OBJECTS; if (e->first_button < 0) e->first_button = s;
        e->has_rect[s] = 1; e->rect[s][0] = e->rect[s][1] = -t->button_rect; e->rect[s][2] = e->rect[s][3] = t->button_rect;
        break;
      case ORC_BOX: case ORC_ROD: case ORC_BALL:
        o->keepout = t->box_keepout; o->group = ORC_GROUP_OBJECTS; e->box_slot = s;
        if (t->box_rect > 0) { e->has_rect[s] = 1; e->rect[s][0] = e->rect[s][1] = -t->box_rect; e->rect[s][2] = e->rect[s][3] = t->box_rect; }
        break;
      default: break;
    }
  }
  if (e->task == ORC_T_HAUL_BOX) e->tendon_slot = e->box_slot;
  /* modify_tree variants: go_to_goal_damping.py:12-17, go_to_goal_motor.py:12-16 */
  e->damp_x = e->damp_y = PT_DAMP_XY; e->damp_z = PT_DAMP_Z; e->gear_x = PT_GEAR_X; e->gear_z = PT_GEAR_Z;
  if (e->task == ORC_T_GO_TO_GOAL_DAMPING) e->damp_x = e->damp_y = PT_DAMP_XY * 0.1;
  if (e->task == ORC_T_GO_TO_GOAL_MOTOR) e->gear_x = PT_GEAR_X * 10.0;
}

orc_env* orc_env_create(int robot, int task, const orc_config* cfg) {
  orc_env* e = (orc_env*)calloc(1, sizeof(orc_env));
  e->robot = robot; e->task = task;
  if (cfg) e->cfg = *cfg; else orc_default_config(&e->cfg);
  e->h = robot == ORC_CAR ? CAR_TIMESTEP : PT_TIMESTEP; e->nsub = robot == ORC_CAR ? CAR_NSUB : PT_NSUB;
  e->cq[0] = 1.0;
  e->ctrl_lo[0] = e->ctrl_lo[1] = -1.0; e->ctrl_hi[0] = e->ctrl_hi[1] = 1.0; /* point.xml:7-8 */
  e->bound = e->cfg.max_bound;                                               /* world.py:78 */
  e->new_task = 1;
  /* task-instance state that survives env.reset() (catch_goal.py:12-18, press_buttons.py:20-24) */
  e->cg_cur = 1.0; e->cg_next = 0.2; e->cg_timer = 0;
  e->btn_state = 1; /* State.NORMAL */ e->btn_timer = BUTTON_DELAY; e->goal_button = 0;
  e->active_mask = 0;
  setup_slots(e);
  if (task == ORC_T_COLLECT) e->active_mask = (1u << e->nbuttons) - 1u; /* collect.py:15-16 */
  return e;
}
void orc_env_destroy(orc_env* e) { free(e); }
void orc_env_seed(orc_env* e, uint64_t seed, uint32_t gid) { e->seed = seed; e->gid = gid; e->replay_mode = 0; }
void orc_env_set_replay(orc_env* e, const double* u, int n) { e->replay = u; e->rn = n; e->rpos = 0; e->replay_mode = 1; }
int orc_env_replay_pos(const orc_env* e) { return e->rpos; }
double orc_env_bound(const orc_env* e) { return e->bound; }
void orc_env_new_task(orc_env* e) { e->new_task = 1; }

/* utils.draw_placement, utils.py:22-25,28-70: constrain by keepout, x drawn before y */
void orc_draw_placement(const double rect[4], double keepout, double u1, double u2, double* xy) {
  double xmin = rect[0] + keepout, ymin = rect[1] + keepout, xmax = rect[2] - keepout, ymax = rect[3] - keepout;
  xy[0] = xmin + (xmax - xmin) * u1; /* rs.uniform(xmin, xmax) */
  xy[1] = ymin + (ymax - ymin) * u2;
}

/* World._sample_layout, world.py:191-217.  Returns 1 on success. */
static int sample_layout(orc_env* e, double* rxy) {
  double px[1 + ORC_MAX_OBJ], py[1 + ORC_MAX_OBJ], pk[1 + ORC_MAX_OBJ];
  int np = 0;
  for (int idx = -1; idx < e->nobj; ++idx) { /* -1 = robot (world.py:83-85 puts it first) */
    double keepout = idx < 0 ? e->cfg.robot_keepout : e->obj[idx].keepout;
    const double* rect = (idx >= 0 && e->has_rect[idx]) ? e->rect[idx] : e->extents;
    int conflicted = 1;
    double xy[2];
    for (int k = 0; k < 1000; ++k) {                        /* :207 */
      if (--e->draws_left < 0) return -1;
      double u1, u2;
      rng_pair(e, 0, &u1, &u2);
      orc_draw_placement(rect, keepout, u1, u2, xy);        /* :209-210 */
      int valid = 1;
      for (int j = 0; j < np; ++j) {                        /* :194-200 */
        double dx = xy[0] - px[j], dy = xy[1] - py[j];
        double dist = sqrt(dx * dx + dy * dy);
        if (dist < pk[j] + e->cfg.placements_margin + keepout) { valid = 0; break; }
      }
      if (valid) { conflicted = 0; break; }
    }
    if (conflicted) return 0;                               /* :214-215 */
    px[np] = xy[0]; py[np] = xy[1]; pk[np] = keepout; ++np;
    if (idx < 0) { rxy[0] = xy[0]; rxy[1] = xy[1]; } else { e->obj[idx].x = xy[0]; e->obj[idx].y = xy[1]; }
  }
  return 1;
}

/* GoToGoal._resample_goal_position, go_to_goal.py:59-80 (rect grows x1.01 after EVERY failed draw) */
static int resample_goal(orc_env* e, int stream) {
  double rect[4] = {-1.5, -1.5, 1.5, 1.5};
  for (long j = 0; j < 500000; ++j) {
    double u1, u2, xy[2];
    rng_pair(e, stream, &u1, &u2);
    orc_draw_placement(rect, GOAL_KEEPOUT, u1, u2, xy);
    int valid = 1;
    { double dx = xy[0] - e->q[0], dy = xy[1] - e->q[1];
      if (sqrt(dx * dx + dy * dy) < e->cfg.robot_keepout + GOAL_KEEPOUT) valid = 0; }
    for (int s = 0; valid && s < e->nobj; ++s) {
      if (s == e->goal_slot) continue;
      double dx = xy[0] - e->obj[s].x, dy = xy[1] - e->obj[s].y;
      if (sqrt(dx * dx + dy * dy) < e->obj[s].keepout + GOAL_KEEPOUT) valid = 0;
    }
    if (valid) { e->obj[e->goal_slot].x = xy[0]; e->obj[e->goal_slot].y = xy[1]; return 0; }
    for (int k = 0; k < 4; ++k) rect[k] = rect[k] * 1.01; /* utils.py:118-119 */
  }
  return 1;
}

static double dist2d(double ax, double ay, double bx, double by) { double dx = ax - bx, dy = ay - by; return sqrt(dx * dx + dy * dy); }

static void update_goal_button(orc_env* e) { /* press_buttons.py:78-91 */
  for (int i = 0; i < e->nbuttons; ++i) {
    orc_obj* b = &e->obj[e->first_button + i];
    if (e->btn_state == 0) b->group = ORC_GROUP_INACTIVE;
    else b->group = (i == e->goal_button) ? ORC_GROUP_GOAL : ORC_GROUP_OBJECTS;
  }
}
static void sample_goal_button(orc_env* e, int stream) { /* press_buttons.py:70-76 */
  double u = rng_single(e, stream);
  int k = (int)(u * e->nbuttons); if (k >= e->nbuttons) k = e->nbuttons - 1;
  e->goal_button = k;
  e->btn_timer = BUTTON_DELAY;
  const orc_obj* b = &e->obj[e->first_button + k];
  e->last_dist[0] = dist2d(e->q[0], e->q[1], b->x, b->y);
}

/* task.reset(): go_to_goal.py:50-57, push_box.py:94-100, press_buttons.py:65-68, collect.py:41-47,
 * catch_goal.py:36-40, unsupervised.py:69-76 */
static int task_reset(orc_env* e, int stream) {
  const task_spec* t = &TASKS[e->task];
  if (t->kind == 1) {
    if (e->task == ORC_T_COLLECT) {
      for (int i = 0; i < e->nbuttons; ++i) e->obj[e->first_button + i].group = ORC_GROUP_GOAL;
      e->active_mask = (1u << e->nbuttons) - 1u;
    } else { sample_goal_button(e, stream); update_goal_button(e); }
    return 0;
  }
  if (resample_goal(e, stream)) return 1;
  const orc_obj* g = &e->obj[e->goal_slot];
  e->last_dist[0] = dist2d(e->q[0], e->q[1], g->x, g->y);   /* go_to_goal.py:54-55 (2-D, quirk D1) */
  if (e->task == ORC_T_CATCH_GOAL) { e->cg_ox = g->x; e->cg_oy = g->y; }
  if (t->kind == 2) {
    const orc_obj* b = &e->obj[e->box_slot];
    e->last_dist[1] = dist2d(g->x, g->y, b->x, b->y);        /* _last_box_goal_distance */
    e->last_dist[0] = dist2d(e->q[0], e->q[1], b->x, b->y);  /* _last_box_distance (overwrites the unused goal dist) */
  }
  return 0;
}

int orc_env_reset(orc_env* e, uint32_t episode) {
  const task_spec* t = &TASKS[e->task];
  e->episode = episode; e->ctr[0] = e->ctr[1] = e->ctr[2] = 0;
  setup_slots(e);
  /* World.__init__, world.py:72-78: Task.ctrl_scale (standard Cauchy per actuator, task.py:85-89) and
   * Task.constraint_bound (U(0, max_bound), task.py:91-94) are drawn once per Task instance and cached on it, so they
   * survive env.reset() and change with env.set_task / reset(options={'task': ...}).  Philox stream 3; Cauchy by
   * inversion tan(pi (u - 1/2)).  Not drawn in replay mode (the golden harness injects the reference's ranges). */
  if (e->new_task) {
    e->new_task = 0;
    if (!e->replay_mode) {
      double u[2];
      orc_philox_uniform2(e->seed, 0u, episode, e->gid, 3u, u);
      for (int k = 0; k < 2; ++k) {
        double sn, cs;
        sag_sincos(PI_D * (u[k] - 0.5), &sn, &cs);
        double sc = (sn / cs) * e->cfg.robot_ctrl_range_scale + 1.0;
        e->ctrl_lo[k] = -1.0 * sc; e->ctrl_hi[k] = 1.0 * sc;         /* mujoco_bridge.py:164-166 */
      }
      orc_philox_uniform2(e->seed, 1u, episode, e->gid, 3u, u);
      e->bound = e->cfg.random_bound ? 0.0 + (e->cfg.max_bound - 0.0) * u[0] : e->cfg.max_bound;
    }
  }
  e->draws_left = e->cfg.max_layout_draws > 0 ? e->cfg.max_layout_draws : (1L << 22);
  /* World._generate_new_layout, world.py:172-189 (10000 attempts; the extents-growth fallback is
   * broken in the reference -- quirk D12 -- and is treated as ResamplingError) */
  double rxy[2] = {0, 0};
  int ok = 0;
  for (int a = 0; a < 10000; ++a) { int r = sample_layout(e, rxy); if (r < 0) break; if (r) { ok = 1; break; } }
  if (!ok) return 1;
  /* World._build_world_config, world.py:108-137: yaw draws */
  e->robot_rot = TWO_PI * rng_single(e, 0);                /* :115, utils.py:11-13 */
  for (int s = 0; s < e->nobj; ++s) {
    int ty = e->obj[s].type;
    if (ty == ORC_HAZARD || ty == ORC_VASE || ty == ORC_GREMLIN || ty == ORC_PILLAR) e->obj[s].yaw = TWO_PI * rng_single(e, 0);
  }
  if (e->task == ORC_T_HAUL_BOX) {                         /* haul_box.py:17-18, quirk D16 */
    e->obj[e->box_slot].x = rxy[0] + BOX_SIZE * 3.0; e->obj[e->box_slot].y = rxy[1];
  }
  if (t->kind == 0 || t->kind == 2) e->obj[e->goal_slot].yaw = TWO_PI * rng_single(e, 0); /* go_to_goal.py:26-27 */
  if (t->kind == 2 && t->box_type == ORC_BOX) e->obj[e->box_slot].yaw = TWO_PI * rng_single(e, 0); /* push_box.py:35 */
  if (t->kind == 1) for (int i = 0; i < e->nbuttons; ++i) e->obj[e->first_button + i].yaw = TWO_PI * rng_single(e, 0); /* press_buttons.py:35 */
  /* MujocoBridge.rebuild: fresh physics, mujoco_bridge.py:170-175 */
  e->q[0] = rxy[0]; e->q[1] = rxy[1]; e->q[2] = e->robot_rot;
  e->v[0] = e->v[1] = e->v[2] = 0.0; e->ctrl[0] = e->ctrl[1] = 0.0;
  e->wheel_w[0] = e->wheel_w[1] = 0.0; e->cq[0] = 1.0; e->cq[1] = e->cq[2] = e->cq[3] = 0.0;
  e->time = 0.0; e->error = 0;
  /* gremlins: weld anchor = spawn pose; the mocap bodies start at the world origin (primitive_objects.py:24-31,71-77) */
  e->mocap_pos[0] = e->mocap_pos[1] = e->mocap_kin[0] = e->mocap_kin[1] = 0.0;
  for (int s = 0; s < e->nobj; ++s) { e->gspawn[s][0] = e->obj[s].x; e->gspawn[s][1] = e->obj[s].y; e->gspawn[s][2] = e->obj[s].yaw; }
  orc_phys_forward(e);
  /* World.reset -> task.reset, world.py:167-170 */
  if (task_reset(e, 0)) return 1;
  orc_phys_forward(e);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
static int robot_touches_slot(const orc_env* e, int slot) { /* mujoco_bridge.robot_contacts([name]), :177-191 */
  int n = 0;
  for (int i = 0; i < e->ncon; ++i) {
    const orc_contact* c = &e->con[i];
    if ((c->sa == -1 && c->sb == slot) || (c->sb == -1 && c->sa == slot)) ++n;
  }
  return n;
}

/* World.compute_cost, world.py:144-155 */
static double compute_cost(orc_env* e) {
  orc_phys_forward(e);                                              /* :145 */
  double cost = 0.0;
  for (int i = 0; i < e->ncon; ++i) {                               /* :146 robot_contacts(OBSTACLES) */
    const orc_contact* c = &e->con[i];
    int other = c->sa == -1 ? c->sb : (c->sb == -1 ? c->sa : -2);
    if (other < 0) continue;
    int ty = e->obj[other].type;
    if (ty == ORC_HAZARD || ty == ORC_VASE || ty == ORC_GREMLIN || ty == ORC_PILLAR) cost += 1.0;
  }
  for (int s = 0; s < e->nobj; ++s) {                               /* :148-153 */
    if (e->obj[s].type != ORC_HAZARD) continue;
    double dist = dist2d(e->q[0], e->q[1], e->obj[s].x, e->obj[s].y);
    if (dist <= e->cfg.hazards_size) cost += 1.0;
  }
  return cost > 0.0 ? 1.0 : 0.0;                                    /* :155 */
}

static double tolerance01(double x, double hi) { return (x >= 0.0 && x <= hi) ? 1.0 : 0.0; } /* dm_control rewards.tolerance, margin 0 [EXT] */

/* task.compute_reward family.  returns 1 on ResamplingError */
static int compute_reward(orc_env* e, double* reward) {
  const task_spec* t = &TASKS[e->task];
  reward[0] = reward[1] = 0.0;
  if (t->kind == 0) {
    /* GoToGoal.compute_reward go_to_goal.py:31-45 (3-D distance: robot z vs goal z, quirk D1) */
    const orc_obj* g = &e->obj[e->goal_slot];
    double dx = e->q[0] - g->x, dy = e->q[1] - g->y, dz = PT_Z - GOAL_Z;
    double distance = sqrt(dx * dx + dy * dy + dz * dz);
    double r = e->last_dist[0] - distance;
    if (e->task == ORC_T_GO_TO_GOAL_SCARCE) r = tolerance01(distance, GOAL_SIZE * 1.5) * r; /* go_to_goal_scarce.py:26-32 */
    e->last_dist[0] = distance;
    if (distance <= GOAL_SIZE) {
      if (task_reset(e, 1)) return 1;
      orc_phys_forward(e);
      r += 1.0;
    }
    if (e->task == ORC_T_UNSUPERVISED) {                   /* unsupervised.py:48-67 */
      double cx = pt_mc() / pt_mass(), cy = 0.0;
      if (e->robot == ORC_CAR) { car_model cm = car_params(); cx = cm.mcx / cm.M; cy = cm.mcy / cm.M; }
      double cs = sag_cos(e->q[2]), sn = sag_sin(e->q[2]);
      double ox = cx * cs - cy * sn, oy = cx * sn + cy * cs;   /* COM offset in the world frame */
      double x = e->q[0] + ox, y = e->q[1] + oy;               /* subtree_com */
      double u = e->v[0] - e->v[2] * oy, v = e->v[1] + e->v[2] * ox; /* subtree_linvel */
      double radius = sqrt(x * x + y * y);
      reward[0] = (((-u * y + v * x) / radius) / (1.0 + fabs(radius - 1.5))) * 1e-1;
      reward[1] = r;
    } else reward[0] = r;
    return 0;
  }
  if (t->kind == 1 && e->task == ORC_T_COLLECT) {          /* collect.py:24-39 */
    if (!e->active_mask) task_reset(e, 1);
    for (int i = 0; i < e->nbuttons; ++i) {
      if (!(e->active_mask & (1u << i))) continue;
      if (robot_touches_slot(e, e->first_button + i)) {
        reward[0] += 1.0;
        e->obj[e->first_button + i].group = ORC_GROUP_INACTIVE;
        e->active_mask &= ~(1u << i);
        break;
      }
    }
    return 0;
  }
  if (t->kind == 1) {                                      /* press_buttons.py:42-63 / press_buttons_scarce.py:22-55 */
    const orc_obj* b = &e->obj[e->first_button + e->goal_button];
    double d = dist2d(e->q[0], e->q[1], b->x, b->y);
    double r = e->task == ORC_T_PRESS_BUTTONS_SCARCE ? 0.0 : e->last_dist[0] - d;
    e->last_dist[0] = d;
    if (robot_touches_slot(e, e->first_button + e->goal_button)) {
      r += 1.0;
      sample_goal_button(e, 1);
      e->btn_state = 0; /* BUTTON_CHANGE */
    }
    if (e->btn_state == 0) {
      if (e->btn_timer != 0) e->btn_timer = e->btn_timer - 1 > 0 ? e->btn_timer - 1 : 0;
      else { e->btn_state = 1; e->btn_timer = BUTTON_DELAY; }
    }
    update_goal_button(e);
    reward[0] = r;
    return 0;
  }
  /* kind 2: PushBox push_box.py:74-92, PushBoxScarce push_box_scarce.py:22-50, HaulBox haul_box.py:34-48 */
  const orc_obj* g = &e->obj[e->goal_slot];
  const orc_obj* b = &e->obj[e->box_slot];
  double r = 0.0;
  if (e->task != ORC_T_HAUL_BOX) {
    double bd = dist2d(e->q[0], e->q[1], b->x, b->y);
    double sh = e->last_dist[0] - bd;
    if (e->task == ORC_T_PUSH_BOX_SCARCE) sh = tolerance01(bd, GOAL_SIZE * 1.70) * sh;
    r += sh;
    e->last_dist[0] = bd;
  }
  double bg = dist2d(b->x, b->y, g->x, g->y);
  r += e->last_dist[1] - bg;
  e->last_dist[1] = bg;
  if (bg <= GOAL_SIZE) {
    if (task_reset(e, 1)) return 1;
    orc_phys_forward(e);
    r += 1.0;
  }
  reward[0] = r;
  return 0;
}

/* CatchGoal.set_mocaps, catch_goal.py:20-31 */
static void set_mocaps(orc_env* e) {
  if (e->cfg.num_gremlins > 0) { /* World.set_mocaps, world.py:157-165: every gremlin's mocap gets the same target */
    double phase = e->time;
    orc_phys_set_mocap_pos(e, sag_sin(phase) * e->cfg.gremlins_travel, sag_cos(phase) * e->cfg.gremlins_travel);
  }
  if (e->task != ORC_T_CATCH_GOAL) return;
  e->cg_timer = e->cg_timer - 1 > 0 ? e->cg_timer - 1 : 0;
  if (e->cg_timer == 0) {
    e->cg_cur = e->cg_next;
    e->cg_next = 0.2 + (1.0 - 0.2) * rng_single(e, 1);
    e->cg_timer = 10;
  }
  double phase = e->time;
  double progress = (10 - e->cg_timer) / 10.0;
  double radius = progress * (e->cg_next - e->cg_cur) + e->cg_cur;
  e->obj[e->goal_slot].x = e->cg_ox + sag_sin(phase) * radius;
  e->obj[e->goal_slot].y = e->cg_oy + sag_cos(phase) * radius;
}

/* SafeAdaptationGym.step, safe_adaptation_gym.py:56-83 */
int orc_env_step(orc_env* e, const double* action, double* obs, double* reward, double* cost, int* done) {
  double a[2] = {action[0], action[1]};
  if (e->replay_mode) {                                    /* :63-65 (recorded rs.normal draws are replayed verbatim) */
    double n0 = rng_single(e, 1), n1 = rng_single(e, 1);
    a[0] += e->cfg.action_noise * n0; a[1] += e->cfg.action_noise * n1;
  } else if (e->cfg.action_noise != 0.0) {
    double u1, u2;
    rng_pair(e, 1, &u1, &u2);
    double rad = sqrt(-2.0 * sag_log(1.0 - u1));               /* Box-Muller */
    a[0] += e->cfg.action_noise * (rad * sag_cos(TWO_PI * u2));
    a[1] += e->cfg.action_noise * (rad * sag_sin(TWO_PI * u2));
  }
  /* :66-67 np.clip = minimum(maximum(a, lo), hi); MuJoCo's own ctrl clamp (x < lo ? lo : x > hi ? hi : x [EXT]) is
   * applied once more where the actuator force is computed (pt_smooth / car_wheel_smooth).  The two differ only for
   * an inverted range, which a Cauchy-scaled ctrlrange can be (lo > hi, world.py:72-73) */
  double u[2];
  for (int k = 0; k < 2; ++k) {
    double t = a[k] > e->ctrl_lo[k] ? a[k] : e->ctrl_lo[k];
    u[k] = t < e->ctrl_hi[k] ? t : e->ctrl_hi[k];
  }
  orc_phys_set_control(e, u);
  set_mocaps(e);                                           /* :71 */
  orc_phys_step(e, e->nsub);                               /* :72 */
  if (e->error) {                                          /* :73-75 */
    orc_env_observation(e, obs);
    reward[0] = -10.0; reward[1] = 0.0; *cost = 0.0; *done = 1;
    return 0;
  }
  orc_phys_forward(e);                                     /* :76 */
  if (compute_reward(e, reward)) return 1;                 /* :77 */
  *cost = compute_cost(e);                                 /* :78 */
  orc_env_observation(e, obs);                             /* :80 */
  *done = 0;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * state access
 * ---------------------------------------------------------------------------------------- */
int orc_env_nobj(const orc_env* e) { return e->nobj; }
void orc_env_get_robot(const orc_env* e, double* o) { for (int k = 0; k < 3; ++k) { o[k] = e->q[k]; o[3 + k] = e->v[k]; } }
void orc_env_set_robot(orc_env* e, const double* in) { for (int k = 0; k < 3; ++k) { e->q[k] = in[k]; e->v[k] = in[3 + k]; } }
void orc_env_get_robot_ext(const orc_env* e, double* o) { o[0] = e->wheel_w[0]; o[1] = e->wheel_w[1]; for (int k = 0; k < 4; ++k) o[2 + k] = e->cq[k]; }
void orc_env_set_robot_ext(orc_env* e, const double* in) { e->wheel_w[0] = in[0]; e->wheel_w[1] = in[1]; for (int k = 0; k < 4; ++k) e->cq[k] = in[2 + k]; }
void orc_env_get_obj(const orc_env* e, int s, orc_obj* out) { *out = e->obj[s]; }
void orc_env_set_obj(orc_env* e, int s, const orc_obj* in) { e->obj[s] = *in; }
void orc_env_get_task_state(const orc_env* e, double* o) {
  o[0] = e->last_dist[0]; o[1] = e->last_dist[1]; o[2] = e->goal_button; o[3] = e->btn_state; o[4] = e->btn_timer;
  o[5] = e->active_mask; o[6] = e->cg_cur; o[7] = e->cg_next; o[8] = e->cg_timer; o[9] = e->cg_ox; o[10] = e->cg_oy;
  o[11] = e->ctr[1]; o[12] = e->time; o[13] = e->robot_rot; o[14] = e->ctrl[0]; o[15] = e->ctrl[1];
}
void orc_env_set_task_state(orc_env* e, const double* o) {
  e->last_dist[0] = o[0]; e->last_dist[1] = o[1]; e->goal_button = (int)o[2]; e->btn_state = (int)o[3]; e->btn_timer = (int)o[4];
  e->active_mask = (unsigned)o[5]; e->cg_cur = o[6]; e->cg_next = o[7]; e->cg_timer = (int)o[8]; e->cg_ox = o[9]; e->cg_oy = o[10];
  e->ctr[1] = (uint32_t)o[11]; e->time = o[12]; e->robot_rot = o[13]; e->ctrl[0] = o[14]; e->ctrl[1] = o[15];
}
void orc_env_set_dyn_params(orc_env* e, double damp_xy, double gear_x) { e->damp_x = e->damp_y = damp_xy; e->gear_x = gear_x; }
void orc_env_get_ctrlrange(const orc_env* e, double* lo, double* hi) { for (int k = 0; k < 2; ++k) { lo[k] = e->ctrl_lo[k]; hi[k] = e->ctrl_hi[k]; } }
void orc_env_set_ctrlrange(orc_env* e, const double* lo, const double* hi) { for (int k = 0; k < 2; ++k) { e->ctrl_lo[k] = lo[k]; e->ctrl_hi[k] = hi[k]; } }

void orc_phys_set_mocap_pos(orc_env* e, double x, double y) { e->mocap_pos[0] = x; e->mocap_pos[1] = y; }
int orc_env_num_gremlins(const orc_env* e) { return e->cfg.num_gremlins; }
void orc_env_slot_types(const orc_env* e, int* types) { task_slot_types_g(e->task, e->cfg.num_gremlins, types); }
void orc_env_get_gremlin_state(const orc_env* e, double* o) {
  o[0] = e->mocap_pos[0]; o[1] = e->mocap_pos[1]; o[2] = e->mocap_kin[0]; o[3] = e->mocap_kin[1];
  int g = 0;
  for (int s = 0; s < e->nobj; ++s) if (e->obj[s].type == ORC_GREMLIN) { for (int k = 0; k < 3; ++k) o[4 + 3 * g + k] = e->gspawn[s][k]; ++g; }
}
void orc_env_set_gremlin_state(orc_env* e, const double* o) {
  e->mocap_pos[0] = o[0]; e->mocap_pos[1] = o[1]; e->mocap_kin[0] = o[2]; e->mocap_kin[1] = o[3];
  int g = 0;
  for (int s = 0; s < e->nobj; ++s) if (e->obj[s].type == ORC_GREMLIN) { for (int k = 0; k < 3; ++k) e->gspawn[s][k] = o[4 + 3 * g + k]; ++g; }
}
void orc_phys_clear(orc_env* e) {
  e->mocap_pos[0] = e->mocap_pos[1] = e->mocap_kin[0] = e->mocap_kin[1] = 0.0;
  e->nobj = 0; e->goal_slot = e->box_slot = e->first_button = -1; e->nbuttons = 0; e->tendon_slot = -1;
  e->time = 0.0; e->error = 0; e->ncon = 0;
  e->v[0] = e->v[1] = e->v[2] = 0.0; e->ctrl[0] = e->ctrl[1] = 0.0;
  e->wheel_w[0] = e->wheel_w[1] = 0.0; e->cq[0] = 1.0; e->cq[1] = e->cq[2] = e->cq[3] = 0.0;
}
int orc_phys_add_obj(orc_env* e, int type, double x, double y, double yaw, double keepout, int group) {
  if (e->nobj >= ORC_MAX_OBJ) return -1;
  int s = e->nobj++;
  orc_obj* o = &e->obj[s];
  memset(o, 0, sizeof(*o));
  o->type = type; o->x = x; o->y = y; o->yaw = yaw; o->keepout = keepout; o->group = group;
  e->gspawn[s][0] = x; e->gspawn[s][1] = y; e->gspawn[s][2] = yaw;
  e->has_rect[s] = 0;
  if (type == ORC_GOAL) e->goal_slot = s;
  if (type == ORC_BOX || type == ORC_ROD || type == ORC_BALL) { e->box_slot = s; if (e->task == ORC_T_HAUL_BOX) e->tendon_slot = s; }
  if (type == ORC_BUTTON) { if (e->first_button < 0) e->first_button = s; e->nbuttons++; }
  return s;
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline helper: synthetic random-action rollout (bench.py cpu_baseline / --impl reference)
 * ---------------------------------------------------------------------------------------- */
typedef struct { orc_env** envs; int lo, hi, steps; double sr, sc; long count; } rollout_job;

static void* rollout_worker(void* arg) {
  rollout_job* j = (rollout_job*)arg;
  for (int i = j->lo; i < j->hi; ++i) {
    orc_env* e = j->envs[i];
    double obs[ORC_OBS_MAX], rew[2], cost, u[2], act[2];
    int done;
    for (int t = 0; t < j->steps; ++t) {
      orc_philox_uniform2(e->seed, e->ctr[2]++, e->episode, e->gid, 2u, u);
      act[0] = 2.0 * u[0] - 1.0; act[1] = 2.0 * u[1] - 1.0;
      if (orc_env_step(e, act, obs, rew, &cost, &done)) break;
      j->sr += rew[0]; j->sc += cost; j->count++;
    }
  }
  return 0;
}

long orc_batch_rollout(orc_env** envs, int n, int steps, int nthreads, double* sum_reward, double* sum_cost) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  rollout_job jobs[256];
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].envs = envs; jobs[t].steps = steps; jobs[t].sr = jobs[t].sc = 0.0; jobs[t].count = 0;
    jobs[t].lo = (int)((long)n * t / nthreads); jobs[t].hi = (int)((long)n * (t + 1) / nthreads);
    pthread_create(&th[t], 0, rollout_worker, &jobs[t]);
  }
  double sr = 0.0, sc = 0.0; long count = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], 0); sr += jobs[t].sr; sc += jobs[t].sc; count += jobs[t].count; }
  *sum_reward = sr; *sum_cost = sc;
  return count;
}

/* deterministic math exports (tests/test_detmath.py): fn 0 sincos, 1 atan2(a,b), 2 log */
void orc_detmath(int fn, const double* a, const double* b, int n, double* out, double* out2) {
  for (int i = 0; i < n; ++i) {
    if (fn == 0) sag_sincos(a[i], out + i, out2 + i);
    else if (fn == 1) out[i] = sag_atan2(a[i], b[i]);
    else out[i] = sag_log(a[i]);
  }
}
