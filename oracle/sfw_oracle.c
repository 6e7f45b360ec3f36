/*
 * sfw_oracle.c — CPU restatement (double, plain C) of the reference's (v,w) scoring path.
 * TEST INFRASTRUCTURE ONLY (see sfw_oracle.h): checker and timed CPU baseline, never the product.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * The lightsfm arithmetic is a restatement of an un-vendored, un-pinned dependency
 * (package.xml:31): parity there is UNPINNED.  Everything else is validated against the
 * reference's own object code (oracle/_ref) by tests/test_oracle_vs_ref.py.
 */
#include "sfw_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif

/* ------------------------------------------------------------------------------------------ */
/* defaults                                                                                    */
/* ------------------------------------------------------------------------------------------ */

static const SfwSfmParams kDefaultSfm = {2.0, 10.0, 0.2, 2.1, 3.0, 2.0, 1.0, 2.0, 0.35, 2.0, 3.0, 0.5};

/* ------------------------------------------------------------------------------------------ */
/* branch probe (sfw_oracle.h: SfwOracleProbe): record the decisions taken within a margin of  */
/* their switching surface, and take listed decisions the other way                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  SfwOracleMargins *mg; /* nullable */
  SfwOracleProbe *pr;   /* nullable */
  int step;
} Ctx;

/* the decision this run must take for (kind, step, a, b), or `natural` when it is not listed */
static int probe_decide(Ctx *cx, int kind, int step, int a, int b, int natural) {
  const SfwOracleProbe *pr = cx ? cx->pr : NULL;
  if (!pr)
    return natural;
  for (uint32_t i = 0; i < pr->n_flips; ++i) {
    const SfwOracleEvent *f = &pr->flips[i];
    if (f->kind == kind && f->step == step && f->a == a && f->b == b)
      return f->decision;
  }
  return natural;
}

static void probe_record(Ctx *cx, int kind, int step, int a, int b, int decision, double margin, double weight) {
  SfwOracleProbe *pr = cx ? cx->pr : NULL;
  if (!pr)
    return;
  uint32_t n = pr->n_events < pr->max_events ? pr->n_events : pr->max_events;
  for (uint32_t i = n; i-- > 0;) { /* both directions of a pair report the same decision once */
    SfwOracleEvent *e = &pr->events[i];
    if (e->step + 1 < step)
      break;
    if (e->step == step && e->kind == kind && e->a == a && e->b == b) {
      if (margin < e->margin)
        e->margin = margin;
      if (weight > e->weight)
        e->weight = weight;
      return;
    }
  }
  if (pr->n_events < pr->max_events) {
    SfwOracleEvent *e = &pr->events[pr->n_events];
    e->kind = kind;
    e->step = step;
    e->a = a;
    e->b = b;
    e->decision = decision;
    e->reserved0 = 0;
    e->margin = margin;
    e->weight = weight;
  }
  pr->n_events++;
}

/* ------------------------------------------------------------------------------------------ */
/* costmap access: nav2 Costmap2D::worldToMap / getCost  [external, SURVEY.md App. C]          */
/* ------------------------------------------------------------------------------------------ */

static double cell_margin_1d(double w, double origin, double res) {
  double c = (w - origin) / res;
  double f = c - floor(c);
  double m = f < 1.0 - f ? f : 1.0 - f;
  return m * res;
}

static int world_to_map(const SfwScene *s, double wx, double wy, unsigned int *mx, unsigned int *my,
                        double *margin) {
  if (wx < s->origin_x || wy < s->origin_y)
    return 0;
  *mx = (unsigned int)((wx - s->origin_x) / s->resolution);
  *my = (unsigned int)((wy - s->origin_y) / s->resolution);
  if (margin) {
    double a = cell_margin_1d(wx, s->origin_x, s->resolution);
    double b = cell_margin_1d(wy, s->origin_y, s->resolution);
    double m = a < b ? a : b;
    if (m < *margin)
      *margin = m;
  }
  return *mx < s->size_x && *my < s->size_y;
}

static unsigned char get_cost(const SfwScene *s, unsigned int mx, unsigned int my) {
  return s->costmap[(size_t)my * s->size_x + mx];
}

/* ------------------------------------------------------------------------------------------ */
/* Bresenham: include/social_force_window_planner/line_iterator.hpp:37-124                     */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  int x, y, curpixel, numpixels;
  int xinc1, xinc2, yinc1, yinc2, den, num, numadd;
} LineIt;

static void line_init(LineIt *l, int x0, int y0, int x1, int y1) {
  int deltax = abs(x1 - x0), deltay = abs(y1 - y0);
  l->x = x0;
  l->y = y0;
  l->curpixel = 0;
  if (x1 >= x0) { /* line_iterator.hpp:43-51 */
    l->xinc1 = 1;
    l->xinc2 = 1;
  } else {
    l->xinc1 = -1;
    l->xinc2 = -1;
  }
  if (y1 >= y0) { /* :53-61 */
    l->yinc1 = 1;
    l->yinc2 = 1;
  } else {
    l->yinc1 = -1;
    l->yinc2 = -1;
  }
  if (deltax >= deltay) { /* :63-71 */
    l->xinc1 = 0;
    l->yinc2 = 0;
    l->den = deltax;
    l->num = deltax / 2;
    l->numadd = deltay;
    l->numpixels = deltax;
  } else { /* :72-80 */
    l->xinc2 = 0;
    l->yinc1 = 0;
    l->den = deltay;
    l->num = deltay / 2;
    l->numadd = deltax;
    l->numpixels = deltay;
  }
}

static int line_valid(const LineIt *l) { return l->curpixel <= l->numpixels; } /* :83 */

static void line_advance(LineIt *l) { /* :85-97 */
  l->num += l->numadd;
  if (l->num >= l->den) {
    l->num -= l->den;
    l->x += l->xinc1;
    l->y += l->yinc1;
  }
  l->x += l->xinc2;
  l->y += l->yinc2;
  l->curpixel++;
}

int sfw_oracle_line_cells(int x0, int y0, int x1, int y1, int *cells_xy, int max_cells) {
  LineIt l;
  int n = 0;
  for (line_init(&l, x0, y0, x1, y1); line_valid(&l); line_advance(&l)) {
    if (cells_xy && n < max_cells) {
      cells_xy[2 * n] = l.x;
      cells_xy[2 * n + 1] = l.y;
    }
    n++;
  }
  return n;
}

/* CostmapModel::pointCost, src/costmap_model.cpp:112-121 (253 is admitted here) */
static double point_cost(const SfwScene *s, int x, int y) {
  unsigned char c = get_cost(s, (unsigned int)x, (unsigned int)y);
  if (c == 255)
    return -2.0;
  if (c == 254)
    return -1.0;
  return (double)c;
}

/* CostmapModel::lineCost, src/costmap_model.cpp:95-110 */
static double line_cost(const SfwScene *s, int x0, int x1, int y0, int y1) {
  double lc = 0.0;
  LineIt l;
  for (line_init(&l, x0, y0, x1, y1); line_valid(&l); line_advance(&l)) {
    double pc = point_cost(s, l.x, l.y);
    if (pc < 0)
      return pc;
    if (lc < pc)
      lc = pc;
  }
  return lc;
}

/* WorldModel::footprintCost(x,y,theta,spec) (world_model.hpp:45-75) followed by
 * CostmapModel::footprintCost(position, footprint) (costmap_model.cpp:21-92). */
double sfw_oracle_footprint_cost(const SfwScene *s, double x, double y, double theta,
                                 double *cell_margin) {
  double cos_th = cos(theta), sin_th = sin(theta); /* world_model.hpp:51-52 */
  unsigned int cell_x, cell_y;
  uint32_t F = s->n_footprint;

  if (!world_to_map(s, x, y, &cell_x, &cell_y, cell_margin)) /* costmap_model.cpp:36-37 */
    return -3.0;

  if (F < 3) { /* costmap_model.cpp:41-48 */
    unsigned char c = get_cost(s, cell_x, cell_y);
    if (c == 255)
      return -2.0;
    if (c == 254 || c == 253)
      return -1.0;
    return (double)c;
  }

  double footprint_cost = 0.0;
  unsigned int x0, y0, x1, y1;
  /* world_model.hpp:54-61: oriented vertex i */
#define VX(i) (x + (s->footprint_xy[2 * (i)] * cos_th - s->footprint_xy[2 * (i) + 1] * sin_th))
#define VY(i) (y + (s->footprint_xy[2 * (i)] * sin_th + s->footprint_xy[2 * (i) + 1] * cos_th))
  for (uint32_t i = 0; i + 1 < F; ++i) { /* costmap_model.cpp:56-72 */
    if (!world_to_map(s, VX(i), VY(i), &x0, &y0, cell_margin))
      return -3.0;
    if (!world_to_map(s, VX(i + 1), VY(i + 1), &x1, &y1, cell_margin))
      return -3.0;
    double lc = line_cost(s, (int)x0, (int)x1, (int)y0, (int)y1);
    footprint_cost = lc > footprint_cost ? lc : footprint_cost;
    if (lc < 0)
      return lc;
  }
  /* closing edge last -> first, costmap_model.cpp:74-87 */
  if (!world_to_map(s, VX(F - 1), VY(F - 1), &x0, &y0, cell_margin))
    return -3.0;
  if (!world_to_map(s, VX(0), VY(0), &x1, &y1, cell_margin))
    return -3.0;
  double lc = line_cost(s, (int)x0, (int)x1, (int)y0, (int)y1);
  footprint_cost = lc > footprint_cost ? lc : footprint_cost;
  if (lc < 0)
    return lc;
#undef VX
#undef VY
  return footprint_cost;
}

/* ------------------------------------------------------------------------------------------ */
/* lightsfm restatement — UNPINNED                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  double px, py, vx, vy;
  double yaw, linvel, angvel;
  double radius, vdes;
  double gx, gy, gr;
  int has_goal, group, id, teleop;
  double fdx, fdy; /* desiredForce  */
  double fox, foy; /* obstacleForce */
  double fsx, fsy; /* socialForce   */
  double fgx, fgy; /* groupForce    */
  double Fx, Fy;   /* globalForce   */
  double ddx, ddy; /* desiredDirection returned by computeDesiredForce */
} OAgent;

static double wrap_pi(double v) { /* utils::Angle: (-pi, pi] */
  while (v <= -M_PI)
    v += 2.0 * M_PI;
  while (v > M_PI)
    v -= 2.0 * M_PI;
  return v;
}

static void normalized2(double x, double y, double *ox, double *oy) {
  double n = sqrt(x * x + y * y);
  if (n > 0.0) {
    *ox = x / n;
    *oy = y / n;
  } else {
    *ox = x;
    *oy = y;
  }
}

/* App. B-3: force on `a` from `b` */
static void pair_force(const SfwSfmParams *P, double apx, double apy, double avx, double avy,
                       double bpx, double bpy, double bvx, double bvy, double *fx, double *fy,
                       double *theta_out, double *mag_out, int force_sgn) {
  double dx = bpx - apx, dy = bpy - apy;
  double ex, ey;
  normalized2(dx, dy, &ex, &ey);
  double vdx = avx - bvx, vdy = avy - bvy;
  double ix = P->lambda * vdx + ex, iy = P->lambda * vdy + ey;
  double L = sqrt(ix * ix + iy * iy);
  double idx = ix / L, idy = iy / L;
  /* interactionDirection.angleTo(diffDirection) = angle(diffDirection) - angle(interactionDirection) */
  double theta = wrap_pi(wrap_pi(atan2(ey, ex)) - wrap_pi(atan2(idy, idx)));
  double B = P->gamma * L;
  double dn = sqrt(dx * dx + dy * dy);
  double a1 = P->n_prime * B * theta, a2 = P->n * B * theta;
  double fv = -exp(-dn / B - a1 * a1);
  int sgn = theta == 0.0 ? 0 : (theta > 0.0 ? 1 : -1);
  if (force_sgn) /* branch probe: the sign this evaluation must use */
    sgn = force_sgn;
  double fa = -(double)sgn * exp(-dn / B - a2 * a2);
  /* forceVelocity = fv * idir ; forceAngle = fa * leftNormal(idir) = fa * (-idy, idx) */
  *fx = P->force_factor_social * (fv * idx + fa * (-idy));
  *fy = P->force_factor_social * (fv * idy + fa * idx);
  if (theta_out)
    *theta_out = theta;
  if (mag_out)
    *mag_out = P->force_factor_social * exp(-dn / B - a2 * a2);
}

void sfw_oracle_pair_force(const SfwSfmParams *sfm, const double me[4], const double other[4],
                           double out_fxy[2], double *theta_out) {
  const SfwSfmParams *P = sfm ? sfm : &kDefaultSfm;
  pair_force(P, me[0], me[1], me[2], me[3], other[0], other[1], other[2], other[3], &out_fxy[0],
             &out_fxy[1], theta_out, NULL, 0);
}

/* App. B-2 */
static void obstacle_force(const SfwSfmParams *P, double px, double py, double radius,
                           const double *obs, uint32_t M, double *fx, double *fy) {
  double sx = 0.0, sy = 0.0;
  if (M == 0) {
    *fx = 0.0;
    *fy = 0.0;
    return;
  }
  for (uint32_t i = 0; i < M; ++i) {
    double dx = px - obs[2 * i], dy = py - obs[2 * i + 1];
    double dist = sqrt(dx * dx + dy * dy) - radius;
    double ux, uy;
    normalized2(dx, dy, &ux, &uy);
    double m = P->force_factor_obstacle * exp(-dist / P->force_sigma_obstacle);
    sx += m * ux;
    sy += m * uy;
  }
  *fx = sx / (double)M;
  *fy = sy / (double)M;
}

void sfw_oracle_obstacle_force(const SfwSfmParams *sfm, double px, double py, double radius,
                               const double *obstacles_xy, uint32_t n_obstacles, double out_fxy[2]) {
  const SfwSfmParams *P = sfm ? sfm : &kDefaultSfm;
  obstacle_force(P, px, py, radius, obstacles_xy, n_obstacles, &out_fxy[0], &out_fxy[1]);
}

/* App. B-1 */
static void desired_force(const SfwSfmParams *P, OAgent *a, int idx, Ctx *cx) {
  a->ddx = 0.0;
  a->ddy = 0.0;
  if (a->has_goal) {
    double dx = a->gx - a->px, dy = a->gy - a->py;
    double n = sqrt(dx * dx + dy * dy);
    int away = n > a->gr;
    if (cx && !a->teleop) {
      double m = fabs(n - a->gr);
      if (cx->mg && m < cx->mg->goal)
        cx->mg->goal = m;
      if (cx->pr) { /* the same decision update_position took after the previous step (same positions) */
        if (m < cx->pr->goal_margin)
          probe_record(cx, SFW_EV_GOAL, cx->step, idx, 0, !away, m, 0.0);
        away = !probe_decide(cx, SFW_EV_GOAL, cx->step, idx, 0, !away);
      }
    }
    if (away) {
      normalized2(dx, dy, &a->ddx, &a->ddy);
      a->fdx = P->force_factor_desired * (a->ddx * a->vdes - a->vx) / P->relaxation_time;
      a->fdy = P->force_factor_desired * (a->ddy * a->vdes - a->vy) / P->relaxation_time;
      return;
    }
  }
  a->fdx = -a->vx / P->relaxation_time;
  a->fdy = -a->vy / P->relaxation_time;
}

/* App. B-4 (only for groupId >= 0 with >= 2 members) */
static void group_force(const SfwSfmParams *P, OAgent *ag, int n, int idx, Ctx *pc) {
  OAgent *a = &ag[idx];
  a->fgx = 0.0;
  a->fgy = 0.0;
  if (a->group < 0)
    return;
  int cnt = 0;
  double cx = 0.0, cy = 0.0;
  for (int i = 0; i < n; ++i)
    if (ag[i].group == a->group) {
      cnt++;
      cx += ag[i].px;
      cy += ag[i].py;
    }
  if (cnt < 2)
    return;
  cx /= (double)cnt;
  cy /= (double)cnt;
  /* gaze */
  double gzx = 0.0, gzy = 0.0;
  double comx = (1.0 / (double)(cnt - 1)) * ((double)cnt * cx - a->px);
  double comy = (1.0 / (double)(cnt - 1)) * ((double)cnt * cy - a->py);
  double rx = comx - a->px, ry = comy - a->py;
  double ep = a->ddx * rx + a->ddy * ry;
  double dn = sqrt(a->ddx * a->ddx + a->ddy * a->ddy), rn = sqrt(rx * rx + ry * ry);
  double com_angle = wrap_pi(acos(ep / (dn * rn)));
  double vision = wrap_pi(90.0 * M_PI / 180.0);
  if (com_angle > vision) { /* false for NaN, as upstream */
    double dd2 = a->ddx * a->ddx + a->ddy * a->ddy;
    double ddist = ep / dd2;
    gzx = ddist * a->ddx * P->force_factor_group_gaze;
    gzy = ddist * a->ddy * P->force_factor_group_gaze;
  }
  /* coherence */
  rx = cx - a->px;
  ry = cy - a->py;
  double dist = sqrt(rx * rx + ry * ry);
  double maxd = ((double)cnt - 1.0) / 2.0;
  double soft = P->force_factor_group_coherence * (tanh(dist - maxd) + 1.0) / 2.0;
  double chx = rx * soft, chy = ry * soft;
  /* repulsion */
  double rpx = 0.0, rpy = 0.0;
  for (int i = 0; i < n; ++i) {
    if (i == idx || ag[i].group != a->group)
      continue;
    double dx = a->px - ag[i].px, dy = a->py - ag[i].py;
    int touching = sqrt(dx * dx + dy * dy) < a->radius + ag[i].radius;
    if (pc) { /* the repulsion term switches on at contact: a discontinuity of the model */
      double m = fabs(sqrt(dx * dx + dy * dy) - (a->radius + ag[i].radius));
      if (pc->mg && m < pc->mg->collision)
        pc->mg->collision = m;
      if (pc->pr) {
        int lo = idx < i ? idx : i, hi = idx < i ? i : idx;
        if (m < pc->pr->collision_margin)
          probe_record(pc, SFW_EV_GROUP, pc->step, lo, hi, touching, m, sqrt(dx * dx + dy * dy));
        touching = probe_decide(pc, SFW_EV_GROUP, pc->step, lo, hi, touching);
      }
    }
    if (touching) {
      rpx += dx;
      rpy += dy;
    }
  }
  rpx *= P->force_factor_group_repulsion;
  rpy *= P->force_factor_group_repulsion;
  a->fgx = gzx + chx + rpx;
  a->fgy = gzy + chy + rpy;
}

/* sfm::SFM.computeForces(std::vector<Agent>&): call site src/sfw_planner.cpp:592 */
static int theta_sign(double th) { return th == 0.0 ? 0 : (th > 0.0 ? 1 : -1); }

/* |theta| and pi - |theta| are the two places where lightsfm's Angle::sign() flips */
static double theta_margin(double th) {
  double a = fabs(th);
  return a < M_PI - a ? a : M_PI - a;
}

static void compute_forces(const SfwSfmParams *P, OAgent *ag, int n, const double *obs, uint32_t M,
                           Ctx *cx) {
  SfwOracleMargins *mg = cx ? cx->mg : NULL;
  for (int i = 0; i < n; ++i) {
    OAgent *a = &ag[i];
    desired_force(P, a, i, cx);
    obstacle_force(P, a->px, a->py, a->radius, obs, M, &a->fox, &a->foy);
    a->fsx = 0.0;
    a->fsy = 0.0;
    for (int j = 0; j < n; ++j) {
      if (j == i)
        continue;
      double fx, fy, th, mag;
      pair_force(P, a->px, a->py, a->vx, a->vy, ag[j].px, ag[j].py, ag[j].vx, ag[j].vy, &fx, &fy,
                 &th, &mag, 0);
      if (cx && cx->pr) { /* theta(i<-j) == theta(j<-i): one decision per unordered pair and step */
        int lo = i < j ? i : j, hi = i < j ? j : i, nat = theta_sign(th);
        if (theta_margin(th) < cx->pr->theta_margin && mag > cx->pr->theta_min_weight)
          probe_record(cx, SFW_EV_THETA, cx->step, lo, hi, nat, theta_margin(th), mag);
        if (cx->pr->n_flips) {
          int want = probe_decide(cx, SFW_EV_THETA, cx->step, lo, hi, nat);
          if (want != nat)
            pair_force(P, a->px, a->py, a->vx, a->vy, ag[j].px, ag[j].py, ag[j].vx, ag[j].vy, &fx, &fy,
                       &th, &mag, want);
        }
      }
      a->fsx += fx;
      a->fsy += fy;
      if (mg && mag > 1e-6 && fabs(th) < mg->theta)
        mg->theta = fabs(th);
    }
    group_force(P, ag, n, i, cx);
    a->Fx = a->fdx + a->fsx + a->fox + a->fgx;
    a->Fy = a->fdy + a->fsy + a->foy + a->fgy;
  }
}

/* sfm::SFM.updatePosition(std::vector<Agent>&, dt): call site src/sfw_planner.cpp:594 (App. B-5) */
static void update_position(OAgent *ag, int n, double dt, Ctx *cx) {
  SfwOracleMargins *mg = cx ? cx->mg : NULL;
  for (int i = 0; i < n; ++i) {
    OAgent *a = &ag[i];
    if (a->teleop) {
      double imd = a->linvel * dt;
      double ang = a->yaw + a->angvel * dt * 0.5;
      a->px += imd * cos(ang);
      a->py += imd * sin(ang);
      a->yaw = wrap_pi(a->yaw + wrap_pi(a->angvel * dt));
      a->vx = a->linvel * cos(a->yaw);
      a->vy = a->linvel * sin(a->yaw);
    } else {
      a->vx += a->Fx * dt;
      a->vy += a->Fy * dt;
      double vn = sqrt(a->vx * a->vx + a->vy * a->vy);
      if (vn > a->vdes) {
        a->vx /= vn; /* velocity.normalize(); velocity *= desiredVelocity */
        a->vy /= vn;
        a->vx *= a->vdes;
        a->vy *= a->vdes;
      }
      a->yaw = wrap_pi(atan2(a->vy, a->vx));
      a->px += a->vx * dt;
      a->py += a->vy * dt;
    }
    if (a->has_goal) {
      double dx = a->gx - a->px, dy = a->gy - a->py;
      double dn = sqrt(dx * dx + dy * dy);
      int reached = dn <= a->gr;
      if (cx && !a->teleop) {
        double m = fabs(dn - a->gr);
        if (mg && m < mg->goal)
          mg->goal = m;
        if (cx->pr) { /* keyed by the step whose force computation sees it */
          if (m < cx->pr->goal_margin)
            probe_record(cx, SFW_EV_GOAL, cx->step + 1, i, 0, reached, m, 0.0);
          reached = probe_decide(cx, SFW_EV_GOAL, cx->step + 1, i, 0, reached);
        }
      }
      if (reached)
        a->has_goal = 0; /* pop_front; cyclicGoals is false for every agent of the reference */
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* planner kinematics: include/social_force_window_planner/sfw_planner.hpp:399-463             */
/* ------------------------------------------------------------------------------------------ */

static double compute_new_velocity(double vg, double vi, double a_max, double dt) { /* :457-463 */
  if ((vg - vi) >= 0)
    return fmin(vg, vi + a_max * dt);
  return fmax(vg, vi - a_max * dt);
}

static float normalize_angle_f(float val, float mn, float mx) { /* :399-407, all-float arithmetic */
  float norm;
  if (val >= mn)
    norm = mn + fmodf((val - mn), (mx - mn));
  else
    norm = mx - fmodf((mn - val), (mx - mn));
  return norm;
}

/* SFWPlanner::scoreTrajectory, src/sfw_planner.cpp:475-676 (+ computeSocialWork :678-705) */
static double score_trajectory(const SfwParams *params, const SfwSfmParams *sfm,
                               const SfwScene *scene, double vx_samp, double vy_samp,
                               double vtheta_samp, double acc_x, double acc_y, double acc_theta,
                               double *pts_xyz, uint32_t max_pts, uint32_t *n_pts,
                               SfwOracleMargins *mg, SfwOracleProbe *pr) {
  const SfwSfmParams *P = sfm ? sfm : &kDefaultSfm;
  Ctx ctx = {mg, pr, 0};
  Ctx *cx = (mg || pr) ? &ctx : NULL;
  if (pr)
    pr->n_events = 0;
  const SfwRobot *R = &scene->robot;
  int n = (int)scene->n_peds + 1;
  OAgent stack_agents[64];
  OAgent *ag = n <= 64 ? stack_agents : (OAgent *)malloc(sizeof(OAgent) * (size_t)n);
  memset(ag, 0, sizeof(OAgent) * (size_t)n);
  if (mg) {
    mg->goal = mg->collision = mg->theta = mg->cell = 1e30;
  }
  if (n_pts)
    *n_pts = 0;

  /* myagents = agents (cpp:486); [0] is the robot as the sensor interface last saw it */
  ag[0].px = R->agent_x;
  ag[0].py = R->agent_y;
  ag[0].vx = R->agent_vx;
  ag[0].vy = R->agent_vy;
  ag[0].radius = R->agent_radius;
  ag[0].teleop = 1;
  ag[0].group = -1;
  ag[0].id = -1;
  ag[0].has_goal = 0;
  for (int j = 1; j < n; ++j) {
    const SfwPed *p = &scene->peds[j - 1];
    ag[j].px = p->x;
    ag[j].py = p->y;
    ag[j].vx = p->vx;
    ag[j].vy = p->vy;
    ag[j].gx = p->goal_x;
    ag[j].gy = p->goal_y;
    ag[j].gr = p->goal_radius;
    ag[j].has_goal = p->has_goal;
    ag[j].vdes = p->desired_velocity;
    ag[j].radius = p->radius;
    ag[j].group = p->group_id;
    ag[j].id = p->id;
  }

  double x_i = R->x, y_i = R->y, theta_i = R->theta;         /* cpp:501-503 */
  double vx_i = R->vx, vy_i = R->vy, vtheta_i = R->vtheta;   /* cpp:505-508 */
  int num_steps = (int)(params->sim_time / params->sim_granularity + 0.5); /* cpp:519 */
  if (num_steps == 0)
    num_steps = 1; /* cpp:523-525 */
  double dt = params->sim_time / num_steps; /* cpp:527 */
  double social_work = 0.0, costmap_cost = 0.0;
  double result = -1.0;
  float rr = params->robot_radius * params->robot_radius; /* float product, cpp:617 */

  for (int i = 0; i < num_steps; ++i) { /* cpp:540 */
    unsigned int cell_x, cell_y;
    ctx.step = i;
    if (!world_to_map(scene, x_i, y_i, &cell_x, &cell_y, NULL)) /* cpp:545-550 */
      goto done;
    double fc = sfw_oracle_footprint_cost(scene, x_i, y_i, theta_i, mg ? &mg->cell : NULL); /* :553 */
    if (fc >= 254.0) /* cpp:555-562 */
      goto done;
    if (fc < 0) /* cpp:565-573 */
      goto done;
    costmap_cost += fc / 255.0; /* cpp:575 */
    if (pts_xyz && n_pts && *n_pts < max_pts) { /* cpp:578 */
      pts_xyz[3 * *n_pts] = x_i;
      pts_xyz[3 * *n_pts + 1] = y_i;
      pts_xyz[3 * *n_pts + 2] = theta_i;
    }
    if (n_pts)
      (*n_pts)++;

    vx_i = compute_new_velocity(vx_samp, vx_i, acc_x, dt); /* cpp:581-583 */
    vy_i = compute_new_velocity(vy_samp, vy_i, acc_y, dt);
    vtheta_i = compute_new_velocity(vtheta_samp, vtheta_i, acc_theta, dt);

    /* cpp:586-588 with hpp:418-446: x and y use the OLD theta and the NEW velocities */
    double nx = x_i + (vx_i * cos(theta_i) + vy_i * cos(M_PI_2 + theta_i)) * dt;
    double ny = y_i + (vx_i * sin(theta_i) + vy_i * sin(M_PI_2 + theta_i)) * dt;
    x_i = nx;
    y_i = ny;
    theta_i = theta_i + vtheta_i * dt;

    compute_forces(P, ag, n, scene->obstacles_xy, scene->n_obstacles, cx); /* cpp:592 */
    double wr = sqrt(ag[0].fsx * ag[0].fsx + ag[0].fsy * ag[0].fsy) +
                sqrt(ag[0].fox * ag[0].fox + ag[0].foy * ag[0].foy); /* cpp:681-682 (values of :592) */
    update_position(ag, n, dt, cx); /* cpp:594 */

    /* cpp:600-610: the robot agent is overwritten with the rolled state */
    ag[0].px = x_i;
    ag[0].py = y_i;
    ag[0].yaw = wrap_pi(theta_i);
    ag[0].linvel = (double)hypotf((float)vx_i, (float)vy_i);
    ag[0].angvel = vtheta_i;
    ag[0].vx = vx_i; /* robot-frame velocity, NOT rotated (cpp:604) */
    ag[0].vy = vy_i;
    ag[0].gx = R->wpx;
    ag[0].gy = R->wpy;
    ag[0].gr = 0.20;
    ag[0].has_goal = 1;

    for (int j = 1; j < n; ++j) { /* cpp:613-627 */
      double dx = ag[0].px - ag[j].px, dy = ag[0].py - ag[j].py;
      double d = dx * dx + dy * dy;
      int hit = d <= (double)rr;
      if (cx) {
        double m = fabs(sqrt(d) - (double)params->robot_radius);
        if (mg && m < mg->collision)
          mg->collision = m;
        if (pr) {
          if (m < pr->collision_margin)
            probe_record(cx, SFW_EV_COLLISION, i, 0, j, hit, m, 0.0);
          hit = probe_decide(cx, SFW_EV_COLLISION, i, 0, j, hit);
        }
      }
      if (hit)
        goto done;
    }

    /* computeSocialWork, cpp:678-705: wr from the forces of :592, wp with the updated states */
    double wp = 0.0;
    for (int j = 1; j < n; ++j) {
      if (ag[j].id == ag[0].id) /* lightsfm (Agent&, vector) overload skips equal ids */
        continue;
      double fx, fy, th, mag;
      /* |f| does not depend on sign(theta): no branch to probe here */
      pair_force(P, ag[j].px, ag[j].py, ag[j].vx, ag[j].vy, ag[0].px, ag[0].py, ag[0].vx, ag[0].vy,
                 &fx, &fy, &th, &mag, 0);
      wp += sqrt(fx * fx + fy * fy);
      if (mg && mag > 1e-6 && fabs(th) < mg->theta)
        mg->theta = fabs(th);
    }
    social_work += wr + wp; /* cpp:629,704 */
  }

  {
    double dx = R->wpx - x_i, dy = R->wpy - y_i; /* cpp:643-644 */
    double d = dx * dx + dy * dy;                /* cpp:647: SQUARED distance */
    double dtheta = atan2(dy, dx);               /* cpp:648 */
    double ang_diff = dtheta - theta_i;          /* cpp:650 */
    ang_diff = (double)normalize_angle_f((float)ang_diff, (float)(-M_PI), (float)M_PI); /* :651 */
    ang_diff = fabs(ang_diff) / M_PI;                                                    /* :652 */
    double vel_diff = fabs(params->max_vel_x - vx_i) / params->max_vel_x;                /* :654 */
    costmap_cost = costmap_cost / num_steps;                                             /* :656 */
    result = (params->vel_weight * vel_diff) + (params->distance_weight * d) +
             (params->angle_weight * ang_diff) + (params->costmap_weight * costmap_cost) +
             (params->social_weight * social_work); /* cpp:663-667 */
  }
done:
  if (ag != stack_agents)
    free(ag);
  return result;
}

double sfw_oracle_score_trajectory(const SfwParams *params, const SfwSfmParams *sfm,
                                   const SfwScene *scene, double vx_samp, double vy_samp,
                                   double vtheta_samp, double acc_x, double acc_y, double acc_theta,
                                   double *pts_xyz, uint32_t max_pts, uint32_t *n_pts,
                                   SfwOracleMargins *mg) {
  return score_trajectory(params, sfm, scene, vx_samp, vy_samp, vtheta_samp, acc_x, acc_y, acc_theta, pts_xyz,
                          max_pts, n_pts, mg, NULL);
}

double sfw_oracle_score_trajectory_probe(const SfwParams *params, const SfwSfmParams *sfm,
                                         const SfwScene *scene, double vx_samp, double vy_samp,
                                         double vtheta_samp, double acc_x, double acc_y, double acc_theta,
                                         SfwOracleProbe *probe) {
  return score_trajectory(params, sfm, scene, vx_samp, vy_samp, vtheta_samp, acc_x, acc_y, acc_theta, NULL, 0,
                          NULL, NULL, probe);
}

/* best-trajectory bookkeeping of findBestAction, src/sfw_planner.cpp:338-344,394-414,426-468 */
void sfw_oracle_argmin(const double *costs, const double *linvels, uint32_t n_v,
                       const double *angvels, uint32_t n_w, SfwBest *best) {
  double best_cost = 10000.0; /* cpp:344 */
  double best_xv = 0.0, best_thetav = 0.0; /* Trajectory() : xv_(0), thetav_(0), trajectory.cpp:16 */
  int found = 0;
  uint32_t best_i = 0, i = 0;
  for (uint32_t a = 0; a < n_v; ++a) {
    for (uint32_t b = 0; b < n_w; ++b, ++i) {
      double linvel = linvels[a], angvel = angvels[b];
      if (linvel == 0.0 && angvel == 0.0) /* cpp:349-352 */
        continue;
      double cost = costs[i];
      if (cost >= 0.0 && cost <= best_cost) { /* cpp:394 */
        if (cost == best_cost && linvel < best_xv) /* cpp:397-401 */
          continue;
        if (cost == best_cost && linvel == best_xv && fabs(angvel) > fabs(best_thetav)) /* :403-407 */
          continue;
        best_cost = cost;
        best_i = i;
        best_xv = linvel;
        best_thetav = angvel;
        found = 1;
      }
    }
  }
  memset(best, 0, sizeof(*best));
  best->valid = found;
  if (found) {
    best->index = best_i;
    best->cost = (float)best_cost;
    best->v = best_xv;
    best->w = best_thetav;
  }
}

static void score_range(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                        const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                        uint32_t begin, uint32_t end, double *costs, SfwOracleMargins *margins) {
  (void)n_v;
  for (uint32_t i = begin; i < end; ++i) {
    double linvel = linvels[i / n_w], angvel = angvels[i % n_w];
    if (linvel == 0.0 && angvel == 0.0) {
      costs[i] = -2.0;
      if (margins) {
        margins[i].goal = margins[i].collision = margins[i].theta = margins[i].cell = 1e30;
      }
      continue;
    }
    /* cpp:356-358: acc_x = max_trans_acc, acc_y = 0, acc_theta = max_rot_acc, vy_samp = 0 */
    costs[i] = sfw_oracle_score_trajectory(params, sfm, scene, linvel, 0.0, angvel,
                                           params->max_trans_acc, 0.0, params->max_rot_acc, NULL, 0,
                                           NULL, margins ? &margins[i] : NULL);
  }
}

int sfw_oracle_score(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                     const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                     double *costs_out, SfwBest *best_out, SfwOracleMargins *margins_out) {
  if (!params || !scene || !linvels || !angvels || !costs_out)
    return SFW_ERR_ARG;
  score_range(params, sfm, scene, linvels, n_v, angvels, n_w, 0, n_v * n_w, costs_out, margins_out);
  if (best_out)
    sfw_oracle_argmin(costs_out, linvels, n_v, angvels, n_w, best_out);
  return SFW_OK;
}

typedef struct {
  const SfwParams *params;
  const SfwSfmParams *sfm;
  const SfwScene *scene;
  const double *linvels, *angvels;
  uint32_t n_v, n_w, total, chunk;
  volatile uint32_t *next;
  double *costs;
} MtJob;

static void *mt_worker(void *arg) {
  MtJob *j = (MtJob *)arg;
  for (;;) {
    uint32_t b = __sync_fetch_and_add(j->next, j->chunk);
    if (b >= j->total)
      break;
    uint32_t e = b + j->chunk < j->total ? b + j->chunk : j->total;
    score_range(j->params, j->sfm, j->scene, j->linvels, j->n_v, j->angvels, j->n_w, b, e, j->costs,
                NULL);
  }
  return NULL;
}

int sfw_oracle_score_mt(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                        const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                        double *costs_out, SfwBest *best_out, int n_threads) {
  if (!params || !scene || !linvels || !angvels || !costs_out)
    return SFW_ERR_ARG;
  if (n_threads < 1)
    n_threads = 1;
  if (n_threads > 256)
    n_threads = 256;
  volatile uint32_t next = 0;
  MtJob job = {params, sfm, scene, linvels, angvels, n_v, n_w, n_v * n_w, 8, &next, costs_out};
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t)
    pthread_create(&th[t], NULL, mt_worker, &job);
  for (int t = 0; t < n_threads; ++t)
    pthread_join(th[t], NULL);
  if (best_out)
    sfw_oracle_argmin(costs_out, linvels, n_v, angvels, n_w, best_out);
  return SFW_OK;
}

/* The sample loop over samples [first, first + count) with the branch probe on: per sample the cost and the
 * decisions taken within the thresholds of `cfg` (cfg->flips are applied to EVERY sample: pass none for a
 * grid).  Samples farmed over n_threads pthreads. */
typedef struct {
  const SfwParams *params;
  const SfwSfmParams *sfm;
  const SfwScene *scene;
  const double *linvels, *angvels;
  uint32_t n_w, first, count;
  const SfwOracleProbe *cfg;
  volatile uint32_t *next;
  double *costs;
  SfwOracleEvent *events;
  uint32_t *n_events;
} ProbeJob;

static void *probe_worker(void *arg) {
  ProbeJob *j = (ProbeJob *)arg;
  for (;;) {
    uint32_t k = __sync_fetch_and_add(j->next, 1u);
    if (k >= j->count)
      break;
    uint32_t i = j->first + k;
    double linvel = j->linvels[i / j->n_w], angvel = j->angvels[i % j->n_w];
    if (linvel == 0.0 && angvel == 0.0) {
      j->costs[k] = -2.0;
      j->n_events[k] = 0;
      continue;
    }
    SfwOracleProbe pr = *j->cfg;
    pr.events = j->events + (size_t)k * j->cfg->max_events;
    pr.n_events = 0;
    j->costs[k] = score_trajectory(j->params, j->sfm, j->scene, linvel, 0.0, angvel, j->params->max_trans_acc,
                                   0.0, j->params->max_rot_acc, NULL, 0, NULL, NULL, &pr);
    j->n_events[k] = pr.n_events;
  }
  return NULL;
}

int sfw_oracle_score_probe(const SfwParams *params, const SfwSfmParams *sfm, const SfwScene *scene,
                           const double *linvels, uint32_t n_v, const double *angvels, uint32_t n_w,
                           uint32_t first, uint32_t count, const SfwOracleProbe *cfg, double *costs_out,
                           SfwOracleEvent *events_out, uint32_t *n_events_out, int n_threads) {
  if (!params || !scene || !linvels || !angvels || !costs_out || !cfg || !events_out || !n_events_out)
    return SFW_ERR_ARG;
  if ((uint64_t)first + count > (uint64_t)n_v * n_w)
    return SFW_ERR_ARG;
  if (n_threads < 1)
    n_threads = 1;
  if (n_threads > 256)
    n_threads = 256;
  volatile uint32_t next = 0;
  ProbeJob job = {params, sfm, scene, linvels, angvels, n_w, first, count, cfg, &next, costs_out, events_out,
                  n_events_out};
  if (n_threads == 1) {
    probe_worker(&job);
    return SFW_OK;
  }
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t)
    pthread_create(&th[t], NULL, probe_worker, &job);
  for (int t = 0; t < n_threads; ++t)
    pthread_join(th[t], NULL);
  return SFW_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* SFMSensorInterface::laserCb — reference src/sensor_interface.cpp:103-229                     */
/* ------------------------------------------------------------------------------------------ */
uint32_t sfw_oracle_laser_obstacles(const SfwLaserScan *scan, float max_obstacle_dist, float person_radius,
                                    double *points_xy_out) {
  uint32_t n = 0;
  float angle = scan->angle_min; /* :118 */
  const double cs = cos(scan->tf_yaw), sn = sin(scan->tf_yaw);
  for (uint32_t i = 0; i < scan->n_ranges; ++i) {
    const float r = scan->ranges[i];
    if (!isnan(r) && isfinite(r) && r < max_obstacle_dist) { /* :120-122 */
      /* math.h is included before use (sensor_interface.hpp:52): cos(float) is the float overload */
      double px = (double)(r * cosf(angle)), py = (double)(r * sinf(angle)); /* :124-125 */
      if (scan->has_tf) { /* :143-170, tf2's transform restricted to the plane */
        const double qx = cs * px - sn * py + scan->tf_x;
        const double qy = sn * px + cs * py + scan->tf_y;
        px = qx;
        py = qy;
      }
      int remove = 0;
      for (uint32_t q = 0; q < scan->n_people; ++q) { /* :210-225 */
        const float dx = (float)(px - scan->people_xy[2 * q]);
        const float dy = (float)(py - scan->people_xy[2 * q + 1]);
        const float d = hypotf(dx, dy);
        if (d <= person_radius) {
          remove = 1;
          break;
        }
      }
      if (!remove) {
        points_xy_out[2 * n] = px;
        points_xy_out[2 * n + 1] = py;
        ++n;
      }
    }
    angle += scan->angle_increment; /* :127 */
  }
  return n;
}
