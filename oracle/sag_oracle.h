/*
 * sag_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, float64, one-environment-at-a-time restatement of the per-step
 * environment loop of lasgroup/safe-adaptation-gym.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may build, load or call this code.  The product path
 * (safe_adaptation_gym_b200/) never links or imports it.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - lidar, cost, rewards, goal/button logic, layout rejection sampling,
 *     yaw draw order, task sampler:  PINNED against golden vectors produced by
 *     running the reference's own Python (tests/golden/make_golden.py imports
 *     /root/reference with dm_control/gym/xmltodict stubbed).
 *   - rigid-body dynamics + contacts (the part the reference delegates to
 *     MuJoCo through dm_control, safe_adaptation_gym.py:72,76): PARITY
 *     UNPINNED.  MuJoCo is not installable in this environment and the
 *     reference holds no numeric test of it.  The model below restates
 *     MuJoCo's published semi-implicit Euler / soft-constraint algorithm for a
 *     planar reduction of point.xml / car.xml (DESIGN.md "Physics model").
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef SAG_ORACLE_H
#define SAG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_OBJ 32
#define ORC_MAX_CON 16      /* contacts per forward pass; more -> PhysicsError (MuJoCo: nconmax) */
#define ORC_MAX_BODIES 8    /* movable bodies with constraint rows per forward pass; more -> PhysicsError */
#define ORC_MAX_PARTS 5
#define ORC_MAX_GREMLINS 4   /* each one adds two weld rows to the solver's row table */
#define ORC_NUM_LIDAR_BINS 16   /* safe_adaptation_gym.py:22 */
#define ORC_OBS_MAX 72          /* car: 48 + 24 */

/* object kinds (primitive_objects.py, tasks/push_box.py) */
enum {
  ORC_NONE = 0, ORC_HAZARD = 1, ORC_VASE = 2, ORC_GREMLIN = 3, ORC_PILLAR = 4,
  ORC_GOAL = 5, ORC_BUTTON = 6, ORC_BOX = 7, ORC_ROD = 8, ORC_BALL = 9
};
/* lidar groups, consts.py:13-16 */
enum { ORC_GROUP_INACTIVE = 0, ORC_GROUP_OBSTACLES = 1, ORC_GROUP_GOAL = 2, ORC_GROUP_OBJECTS = 3 };
/* robots */
enum { ORC_POINT = 0, ORC_CAR = 1 };
/* task ids = alphabetical registry order, benchmark/__init__.py:16-20 */
enum {
  ORC_T_CATCH_GOAL = 0, ORC_T_COLLECT, ORC_T_DRIBBLE_BALL, ORC_T_GO_TO_GOAL,
  ORC_T_GO_TO_GOAL_DAMPING, ORC_T_GO_TO_GOAL_MOTOR, ORC_T_GO_TO_GOAL_SCARCE,
  ORC_T_HAUL_BOX, ORC_T_PRESS_BUTTONS, ORC_T_PRESS_BUTTONS_SCARCE, ORC_T_PUSH_BOX,
  ORC_T_PUSH_BOX_SCARCE, ORC_T_ROLL_ROD, ORC_T_UNSUPERVISED, ORC_NUM_TASKS
};

/* World.DEFAULT, world.py:17-34 */
typedef struct {
  double placements_margin, robot_keepout;
  double hazards_size, vases_size, pillars_size, gremlins_size;
  double hazards_keepout, gremlins_keepout, vases_keepout, pillars_keepout;
  double gremlins_travel;
  double robot_ctrl_range_scale, action_noise, max_bound;
  int random_bound;
  int max_layout_draws; /* draw budget per reset; 0 = default (1<<22). Documented cap. */
  int num_gremlins;     /* Task.obstacles[2] of a user-defined task (task.py:70); 0 in every shipped task */
} orc_config;

typedef struct {
  int type;            /* ORC_* kind */
  int group;           /* lidar group (geom user), may change per step (buttons) */
  double x, y, yaw;    /* planar pose */
  double vx, vy, w;    /* planar velocity (movable kinds only) */
  double keepout;
} orc_obj;

typedef struct {
  int ba, bb;          /* bodies: -1 static, 0 robot, 1+slot movable object */
  int sa, sb;          /* slots for naming: -1 robot, else object slot */
  int ga, gb;          /* geom part index inside body */
  double nx, ny, px, py, dist;
} orc_contact;

typedef struct orc_env orc_env;

void orc_default_config(orc_config* c);

orc_env* orc_env_create(int robot, int task, const orc_config* cfg);
void orc_env_destroy(orc_env* e);

/* RNG: Philox4x32-10 counter mode (matches the GPU) or replay of a uniform stream
 * recorded from the reference's numpy RandomState (golden pinning). */
void orc_env_seed(orc_env* e, uint64_t seed, uint32_t env_gid);
void orc_env_set_replay(orc_env* e, const double* u, int n);
int orc_env_replay_pos(const orc_env* e);

/* reset: safe_adaptation_gym.py:170-172 (_build_world).  returns 0 ok, 1 ResamplingError */
int orc_env_reset(orc_env* e, uint32_t episode);
/* step: safe_adaptation_gym.py:56-83. obs: 60 (point) / 72 (car) doubles; reward[2]
 * (reward[1] only for unsupervised); returns 0 ok, 1 ResamplingError (goal) */
int orc_env_step(orc_env* e, const double* action, double* obs, double* reward, double* cost, int* done);
void orc_env_observation(orc_env* e, double* obs);
int orc_env_obs_dim(const orc_env* e);

/* state access (parity injection) */
int orc_env_nobj(const orc_env* e);
void orc_env_get_robot(const orc_env* e, double* out6);
void orc_env_set_robot(orc_env* e, const double* in6);
/* car extras: wheel rates (2), castor quaternion (4) */
void orc_env_get_robot_ext(const orc_env* e, double* out6);
void orc_env_set_robot_ext(orc_env* e, const double* in6);
void orc_env_get_obj(const orc_env* e, int slot, orc_obj* out);
void orc_env_set_obj(orc_env* e, int slot, const orc_obj* in);
/* task scalars: [last_dist0,last_dist1, goal_button, btn_state, btn_timer, active_mask,
 *                cg_cur, cg_next, cg_timer, cg_ox, cg_oy, step_ctr, time, robot_rot] */
void orc_env_get_task_state(const orc_env* e, double* out16);
void orc_env_set_task_state(orc_env* e, const double* in16);
void orc_env_set_dyn_params(orc_env* e, double damp_xy, double gear_x);
void orc_env_set_ctrlrange(orc_env* e, const double* lo2, const double* hi2);
double orc_env_bound(const orc_env* e);
/* the next orc_env_reset builds the World of a fresh Task instance: redraws ctrl-range scale and constraint bound
 * (world.py:72-78); set at creation */
void orc_env_new_task(orc_env* e);
void orc_env_get_ctrlrange(const orc_env* e, double* lo2, double* hi2);

/* physics-level API: the stand-in for what MujocoBridge exposes (mujoco_bridge.py) */
void orc_phys_set_control(orc_env* e, const double* u2);   /* mujoco_bridge.py:239-241 */
void orc_phys_step(orc_env* e, int nstep);                 /* physics.step(nstep), safe_adaptation_gym.py:72 */
void orc_phys_forward(orc_env* e);                         /* physics.forward(), :76 */
int orc_phys_ncon(const orc_env* e);
void orc_phys_get_contact(const orc_env* e, int i, orc_contact* out);
void orc_phys_sensors(const orc_env* e, double* out);      /* _sensors order, :225-237 */
int orc_phys_error(const orc_env* e);
double orc_phys_time(const orc_env* e);
/* injected-world building for the golden harness (MujocoBridge.rebuild stand-in) */
/* gremlin mocap bodies (primitive_objects.py:57-86): data.mocap_pos of every '<gremlin>mocap' body (world.py:157-165
 * writes the same value to all of them), mujoco_bridge.py:232-233; and the weld anchor = the gremlin's spawn pose */
void orc_phys_set_mocap_pos(orc_env* e, double x, double y);
void orc_env_get_gremlin_state(const orc_env* e, double* out /* mocap_pos[2], mocap_kin[2], then x0,y0,yaw0 per gremlin slot */);
void orc_env_set_gremlin_state(orc_env* e, const double* in);
int orc_env_num_gremlins(const orc_env* e);
void orc_env_slot_types(const orc_env* e, int* types);
void orc_phys_clear(orc_env* e);
int orc_phys_add_obj(orc_env* e, int type, double x, double y, double yaw, double keepout, int group);

/* pure functions (known-answer tests) */
void orc_lidar(double rx, double ry, double ryaw, int n, const double* xs, const double* ys, double* out16);
/* the line-by-line form of safe_adaptation_gym.py:204-223 (full atan2); orc_lidar evaluates the folded form */
void orc_lidar_literal(double rx, double ry, double ryaw, int n, const double* xs, const double* ys, double* out16);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_philox_uniform2(uint64_t seed, uint32_t ctr, uint32_t episode, uint32_t gid, uint32_t stream, double* u2);
void orc_draw_placement(const double rect[4], double keepout, double u1, double u2, double* xy);
/* geometry KATs: returns #contacts, out rows of (nx,ny,px,py,dist) */
int orc_collide_circle_circle(double ax, double ay, double ra, double bx, double by, double rb, double* out5);
int orc_collide_circle_box(double cx, double cy, double r, double bx, double by, double byaw, double hx, double hy,
                           int circle_is_a, double* out5);
int orc_collide_box_box(double ax, double ay, double ayaw, double ahx, double ahy, double bx, double by, double byaw,
                        double bhx, double bhy, double* out10);

void orc_detmath(int fn, const double* a, const double* b, int n, double* out, double* out2);

/* task table (shared facts: obstacle counts etc.) */
int orc_task_nobj(int task);
void orc_task_slot_types(int task, int* types32);

/* multi-env helper for the CPU baseline: steps n envs `steps` times with Philox actions
 * (stream 2), OpenMP over envs; returns env-steps executed. */
long orc_batch_rollout(orc_env** envs, int n, int steps, int nthreads, double* sum_reward, double* sum_cost);

#ifdef __cplusplus
}
#endif
#endif
