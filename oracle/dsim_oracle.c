/*
 * dsim_oracle.c — CPU FP64 ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See dsim_oracle.h.
 *
 * Plain C restatement of the reference hot path:
 *   model      environments/env_gen.py:7-133  (+ MuJoCo compiler `inertiafromgeom`)
 *   physics    mujoco.mj_step, called at environments/mujoco_vecenv.py:404-413   [PARITY UNPINNED]
 *   states     environments/BaseDroneEnv.py:357-380
 *   term.      environments/BaseDroneEnv.py:12-16
 *   rewards    environments/rewards.py:5-368
 *   obs        environments/observation_wrappers.py:7-529
 *   rotations  environments/transformation.py:5-29 (scipy Rotation, Bernardes&Viollet as_euler)
 * The dynamics are written GENERICALLY (spatial vectors about the root subtree COM, world orientation,
 * composite-rigid-body mass matrix, recursive Newton-Euler bias) the way MuJoCo organises them, so that
 * the hand-specialised body-frame CUDA kernel is an independent derivation.
 */
#define _GNU_SOURCE
#include "dsim_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846
#define MJMINVAL 1e-15

/* ------------------------------------------------------------------ small linear algebra */
static void v3_copy(double *r, const double *a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static void v3_zero(double *r) { r[0] = r[1] = r[2] = 0.0; }
static void v3_add(double *r, const double *a, const double *b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static void v3_sub(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static void v3_addto(double *r, const double *a) { r[0] += a[0]; r[1] += a[1]; r[2] += a[2]; }
static void v3_addscl(double *r, const double *a, double s) { r[0] += s * a[0]; r[1] += s * a[1]; r[2] += s * a[2]; }
static void v3_scl(double *r, const double *a, double s) { r[0] = s * a[0]; r[1] = s * a[1]; r[2] = s * a[2]; }
static double v3_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v3_cross(double *r, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
/* r = M v, M row-major 3x3 */
static void m3_mulv(double *r, const double *M, const double *v) {
    double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void m3_tmulv(double *r, const double *M, const double *v) {
    double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void m3_mul(double *r, const double *A, const double *B) {
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(r, t, sizeof t);
}
static void m3_transpose(double *r, const double *A) {
    double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
    memcpy(r, t, sizeof t);
}
/* quaternions are (w,x,y,z) */
static void q_normalize(double *q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < MJMINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void q_mul(double *r, const double *a, const double *b) {
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void q_to_mat(double *R, const double *q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z);           R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);           R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);           R[7] = 2 * (y * z + w * x);           R[8] = w * w - x * x - y * y + z * z;
}
static void q_axis_angle(double *q, const double *axis, double angle) {
    double s = sin(0.5 * angle);
    q[0] = cos(0.5 * angle); q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}

/* ------------------------------------------------------------------ model compiler */
/* dm_control's mjcf `to_xml_string(precision=5)` prints every float attribute with "%.5g"
 * (env_gen.py:128-133); MuJoCo then parses the decimal text back with strtod. */
double orc_round_prec5(double x) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.5g", x);
    return strtod(buf, NULL);
}
static double rnd(double x, int on) { return on ? orc_round_prec5(x) : x; }

/* symmetric 3x3 eigen-decomposition by cyclic Jacobi; eigenvalues sorted descending (MuJoCo's mju_eig3
 * convention); columns of V are eigenvectors, det(V)=+1 */
static void eig3_sym(const double A[9], double eval[3], double V[9]) {
    double a[9];
    memcpy(a, A, sizeof a);
    double v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = fabs(a[1]) + fabs(a[2]) + fabs(a[5]);
        if (off < 1e-300) break;
        int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
        int rotated = 0;
        for (int k = 0; k < 3; k++) {
            int p = P[k], q = Q[k];
            double apq = a[3 * p + q];
            double app = a[3 * p + p], aqq = a[3 * q + q];
            /* MuJoCo's mju_eig3 stops on an ABSOLUTE off-diagonal threshold eigEPS = 1e-12 */
            if (fabs(apq) < 1e-12) continue;
            rotated = 1;
            double theta = (aqq - app) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int r = 0; r < 3; r++) { /* a = a * G */
                double arp = a[3 * r + p], arq = a[3 * r + q];
                a[3 * r + p] = c * arp - s * arq;
                a[3 * r + q] = s * arp + c * arq;
            }
            for (int r = 0; r < 3; r++) { /* a = G^T * a */
                double apr = a[3 * p + r], aqr = a[3 * q + r];
                a[3 * p + r] = c * apr - s * aqr;
                a[3 * q + r] = s * apr + c * aqr;
            }
            for (int r = 0; r < 3; r++) {
                double vrp = v[3 * r + p], vrq = v[3 * r + q];
                v[3 * r + p] = c * vrp - s * vrq;
                v[3 * r + q] = s * vrp + c * vrq;
            }
        }
        if (!rotated) break;
    }
    int idx[3] = {0, 1, 2};
    double d[3] = {a[0], a[4], a[8]};
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (d[idx[j]] > d[idx[i]]) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    for (int k = 0; k < 3; k++) {
        eval[k] = d[idx[k]];
        for (int r = 0; r < 3; r++) V[3 * r + k] = v[3 * r + idx[k]];
    }
    double det = V[0] * (V[4] * V[8] - V[5] * V[7]) - V[1] * (V[3] * V[8] - V[5] * V[6]) + V[2] * (V[3] * V[7] - V[4] * V[6]);
    if (det < 0) for (int r = 0; r < 3; r++) V[3 * r + 2] = -V[3 * r + 2];
}
static void mat_to_quat(double *q, const double *R) {
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
    } else {
        double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
    }
    q_normalize(q);
}

typedef struct { int type; /*0 box 1 cylinder 2 sphere*/ double size[3], pos[3], yaw, mass; } OGeom;

/* MuJoCo compiler, `inertiafromgeom`: geoms with explicit mass -> body mass, COM, principal inertia */
static void body_from_geoms(const OGeom *g, int ng, double *mass, double ipos[3], double iquat[4], double inertia[3]) {
    double M = 0, com[3] = {0, 0, 0};
    for (int i = 0; i < ng; i++) { M += g[i].mass; v3_addscl(com, g[i].pos, g[i].mass); }
    if (M < MJMINVAL) { *mass = 0; v3_zero(ipos); iquat[0] = 1; iquat[1] = iquat[2] = iquat[3] = 0; v3_zero(inertia); return; }
    v3_scl(com, com, 1.0 / M);
    double I[9] = {0};
    for (int i = 0; i < ng; i++) {
        double m = g[i].mass, d[3] = {0, 0, 0};
        if (m <= 0) continue;
        const double *s = g[i].size;
        if (g[i].type == 0) { d[0] = m * (s[1] * s[1] + s[2] * s[2]) / 3; d[1] = m * (s[0] * s[0] + s[2] * s[2]) / 3; d[2] = m * (s[0] * s[0] + s[1] * s[1]) / 3; }
        else if (g[i].type == 1) { double h = 2 * s[1]; d[0] = d[1] = m * (3 * s[0] * s[0] + h * h) / 12; d[2] = m * s[0] * s[0] / 2; }
        else { d[0] = d[1] = d[2] = 0.4 * m * s[0] * s[0]; }
        double c = cos(g[i].yaw), sn = sin(g[i].yaw);
        double Rg[9] = {c, -sn, 0, sn, c, 0, 0, 0, 1}, D[9] = {d[0], 0, 0, 0, d[1], 0, 0, 0, d[2]}, RgT[9], T[9];
        m3_transpose(RgT, Rg); m3_mul(T, Rg, D); m3_mul(T, T, RgT);
        double r[3]; v3_sub(r, g[i].pos, com);
        double r2 = v3_dot(r, r);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) I[3 * a + b] += T[3 * a + b] + m * ((a == b ? r2 : 0.0) - r[a] * r[b]);
    }
    double V[9];
    eig3_sym(I, inertia, V);
    mat_to_quat(iquat, V);
    *mass = M; v3_copy(ipos, com);
}

static void add_geom(OrcModel *m, int body, const OGeom *g) {
    OrcGeom *o = &m->geom[m->ngeom++];
    o->body = body; o->type = g->type; o->yaw = g->yaw;
    memcpy(o->size, g->size, sizeof o->size); memcpy(o->pos, g->pos, sizeof o->pos);
}
static void set_invweight0(OrcModel *m);

void orc_compile(const double params[6], int pendulum_enabled, double frequency, int rp, OrcModel *m) {
    memset(m, 0, sizeof *m);
    memcpy(m->params, params, 6 * sizeof(double));
    double mass = params[0], arm_len = params[1], motor_force = params[2], motor_tau = params[3];
    double pend_len = params[4], weight_mass = params[5];
    int pendulum = pendulum_enabled && pend_len > 0 && weight_mass > 0;   /* env_gen.py:33-36 */
    const double hb = 0.05;                                                /* env_gen.py:38 */
    m->pendulum = pendulum;
    m->nbody = pendulum ? 5 : 3; m->nv = pendulum ? 8 : 6; m->nq = pendulum ? 9 : 7;
    m->timestep = rnd(1.0 / frequency, rp);                                /* env_gen.py:82 */
    m->density = 1.2; m->viscosity = 0.00002;                              /* env_gen.py:83-84 */
    m->gravity[2] = -9.81;
    for (int b = 0; b < ORC_MAXBODY; b++) { m->quat[b][0] = 1; m->iquat[b][0] = 1; m->dofadr[b] = -1; m->qposadr[b] = -1; }
    m->parent[1] = 0; m->parent[2] = 1; m->parent[3] = 2; m->parent[4] = 3;
    /* body 1 = attachment frame with <freejoint> (env_gen.py:123-124): massless */
    m->jnt_type[1] = 1; m->dofadr[1] = 0; m->qposadr[1] = 0;
    for (int d = 0; d < 6; d++) { m->dof_body[d] = 1; m->damping[d] = 0.0; }
    /* body 2 = core_body (env_gen.py:45-64) */
    OGeom g[9]; int ng = 0;
    memset(g, 0, sizeof g);
    g[ng].type = 0; g[ng].size[0] = rnd(hb, rp); g[ng].size[1] = rnd(hb, rp); g[ng].size[2] = rnd(hb / 3, rp);
    g[ng].mass = rnd(0.56 * mass, rp); ng++;
    for (int i = 0; i < 4; i++) {
        double theta = i * PI / 2 - PI / 4;
        double ra = sqrt(2.0) * hb + 0.5 * arm_len, rr = sqrt(2.0) * hb + arm_len;
        g[ng].type = 0; g[ng].size[0] = rnd(arm_len / 2, rp); g[ng].size[1] = rnd(arm_len / 20, rp); g[ng].size[2] = rnd(arm_len / 20, rp);
        g[ng].pos[0] = rnd(ra * cos(theta), rp); g[ng].pos[1] = rnd(ra * sin(theta), rp); g[ng].pos[2] = 0;
        g[ng].yaw = rnd(theta, rp); g[ng].mass = rnd(0.07 * mass, rp); ng++;
        g[ng].type = 1; g[ng].size[0] = 0.01; g[ng].size[1] = 0.01;
        g[ng].pos[0] = rnd(rr * cos(theta) + 0.0, rp); g[ng].pos[1] = rnd(rr * sin(theta) + 0.0, rp); g[ng].pos[2] = rnd(0.0 + 0.015, rp);
        g[ng].mass = rnd(0.04 * mass, rp); ng++;
        m->site_pos[i][0] = rnd(rr * cos(theta), rp); m->site_pos[i][1] = rnd(rr * sin(theta), rp); m->site_pos[i][2] = 0;
        m->gear[i][2] = rnd(motor_force, rp);
        m->gear[i][5] = rnd(motor_force / 100 * ((i % 2) ? -1.0 : 1.0), rp);
        m->tau[i] = rnd(motor_tau, rp);
    }
    m->site_pos[4][2] = rnd(-hb / 4, rp);                                   /* 'sense' site, env_gen.py:48 */
    body_from_geoms(g, ng, &m->mass[2], m->ipos[2], m->iquat[2], m->inertia[2]);
    /* collision geoms of core_body: the 9 above, the massless 'front' box (:47) and the four propeller discs (:60) */
    for (int i = 0; i < ng; i++) add_geom(m, 2, &g[i]);
    {
        OGeom f; memset(&f, 0, sizeof f);
        f.type = 0; f.size[0] = rnd(hb / 3, rp); f.size[1] = rnd(0.15 * hb, rp); f.size[2] = rnd(0.15 * hb, rp); f.pos[0] = rnd(hb + hb / 3, rp);
        add_geom(m, 2, &f);
        for (int i = 0; i < 4; i++) {
            double theta = i * PI / 2 - PI / 4, rr = sqrt(2.0) * hb + arm_len;
            OGeom q; memset(&q, 0, sizeof q);
            q.type = 1; q.size[0] = rnd(arm_len / 1.5, rp); q.size[1] = 0.0025;
            q.pos[0] = rnd(rr * cos(theta) + 0.0, rp); q.pos[1] = rnd(rr * sin(theta) + 0.0, rp); q.pos[2] = rnd(0.0 + 0.025, rp);
            add_geom(m, 2, &q);
        }
    }
    if (pendulum) {
        /* body 3 = link: hinge x, sphere r=.02 m=.01 (env_gen.py:66-68) */
        m->pos[3][2] = rnd(-hb / 2, rp);
        m->jnt_type[3] = 2; m->jnt_axis[3][0] = 1; m->dofadr[3] = 6; m->qposadr[3] = 7; m->dof_body[6] = 3; m->damping[6] = 0.15;
        OGeom s; memset(&s, 0, sizeof s); s.type = 2; s.size[0] = 0.02; s.mass = 0.01;
        body_from_geoms(&s, 1, &m->mass[3], m->ipos[3], m->iquat[3], m->inertia[3]);
        add_geom(m, 3, &s);
        /* body 4 = pendulum: hinge y, pole cylinder + weight box (env_gen.py:69-72) */
        m->jnt_type[4] = 2; m->jnt_axis[4][1] = 1; m->dofadr[4] = 7; m->qposadr[4] = 8; m->dof_body[7] = 4; m->damping[7] = 0.15;
        OGeom p[2]; memset(p, 0, sizeof p);
        p[0].type = 1; p[0].size[0] = 0.005; p[0].size[1] = rnd(pend_len / 2, rp); p[0].pos[2] = rnd(-pend_len / 2, rp); p[0].mass = rnd(0.2 * pend_len, rp);
        double sz = rnd(0.1 * cbrt(weight_mass), rp);
        p[1].type = 0; p[1].size[0] = p[1].size[1] = p[1].size[2] = sz; p[1].pos[2] = rnd(-pend_len, rp); p[1].mass = rnd(weight_mass, rp);
        body_from_geoms(p, 2, &m->mass[4], m->ipos[4], m->iquat[4], m->inertia[4]);
        add_geom(m, 4, &p[0]); add_geom(m, 4, &p[1]);
    }
    set_invweight0(m);
}

/* ------------------------------------------------------------------ dynamics (generic, MuJoCo layout) */
typedef struct {
    double xpos[ORC_MAXBODY][3], xquat[ORC_MAXBODY][4], xmat[ORC_MAXBODY][9];
    double xipos[ORC_MAXBODY][3], ximat[ORC_MAXBODY][9];
    double xanchor[ORC_MAXBODY][3], xaxis[ORC_MAXBODY][3];
    double com[3];                       /* subtree_com of the root (body 1) */
    double cinert[ORC_MAXBODY][10];      /* Ixx Iyy Izz Ixy Ixz Iyz mdx mdy mdz m — about `com`, world axes */
    double cdof[ORC_MAXNV][6];           /* [rot; lin] */
    double cdof_dot[ORC_MAXNV][6];
    double cvel[ORC_MAXBODY][6];
    double M[ORC_MAXNV * ORC_MAXNV];
    double qfrc_bias[ORC_MAXNV], qfrc_passive[ORC_MAXNV], qfrc_actuator[ORC_MAXNV], qfrc_smooth[ORC_MAXNV];
    double site_xpos[5][3];
    double qfrc_constraint[ORC_MAXNV];
    int ncon;
} OData;

static void inert_mulvec(double *res, const double *I, const double *v) {
    /* spatial inertia (about ref point) times motion vector -> [torque; force] */
    const double *w = v, *l = v + 3, *h = I + 6;
    double m = I[9];
    res[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2];
    res[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2];
    res[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2];
    double hx[3]; v3_cross(hx, h, l); v3_addto(res, hx);
    double hw[3]; v3_cross(hw, h, w);
    res[3] = m * l[0] - hw[0]; res[4] = m * l[1] - hw[1]; res[5] = m * l[2] - hw[2];
}
static void cross_motion(double *res, const double *vel, const double *v) {
    double a[3], b[3], c[3];
    v3_cross(a, vel, v); v3_cross(b, vel, v + 3); v3_cross(c, vel + 3, v);
    res[0] = a[0]; res[1] = a[1]; res[2] = a[2];
    res[3] = b[0] + c[0]; res[4] = b[1] + c[1]; res[5] = b[2] + c[2];
}
static void cross_force(double *res, const double *vel, const double *f) {
    double a[3], b[3], c[3];
    v3_cross(a, vel, f); v3_cross(b, vel + 3, f + 3); v3_cross(c, vel, f + 3);
    res[0] = a[0] + b[0]; res[1] = a[1] + b[1]; res[2] = a[2] + b[2];
    res[3] = c[0]; res[4] = c[1]; res[5] = c[2];
}
static double dot6(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }

/* mj_kinematics + mj_comPos */
static void kin_com(const OrcModel *m, const double *qpos, OData *d) {
    memset(d, 0, sizeof *d);
    d->xquat[0][0] = 1; q_to_mat(d->xmat[0], d->xquat[0]);
    for (int i = 1; i < m->nbody; i++) {
        int p = m->parent[i];
        if (m->jnt_type[i] == 1) {
            v3_copy(d->xpos[i], qpos);
            memcpy(d->xquat[i], qpos + 3, 4 * sizeof(double));
            q_normalize(d->xquat[i]);
            v3_copy(d->xanchor[i], d->xpos[i]);
        } else {
            double t[3]; m3_mulv(t, d->xmat[p], m->pos[i]); v3_add(d->xpos[i], d->xpos[p], t);
            q_mul(d->xquat[i], d->xquat[p], m->quat[i]);
            if (m->jnt_type[i] == 2) {
                double R0[9]; q_to_mat(R0, d->xquat[i]);
                v3_copy(d->xanchor[i], d->xpos[i]);                  /* jnt_pos = 0 */
                m3_mulv(d->xaxis[i], R0, m->jnt_axis[i]);
                double ql[4]; q_axis_angle(ql, m->jnt_axis[i], qpos[m->qposadr[i]]);
                q_mul(d->xquat[i], d->xquat[i], ql);
            }
            q_normalize(d->xquat[i]);
        }
        q_to_mat(d->xmat[i], d->xquat[i]);
        double t[3]; m3_mulv(t, d->xmat[i], m->ipos[i]); v3_add(d->xipos[i], d->xpos[i], t);
        double qi[4]; q_mul(qi, d->xquat[i], m->iquat[i]); q_to_mat(d->ximat[i], qi);
    }
    for (int s = 0; s < 5; s++) { double t[3]; m3_mulv(t, d->xmat[2], m->site_pos[s]); v3_add(d->site_xpos[s], d->xpos[2], t); }
    double M = 0; v3_zero(d->com);
    for (int i = 1; i < m->nbody; i++) { M += m->mass[i]; v3_addscl(d->com, d->xipos[i], m->mass[i]); }
    v3_scl(d->com, d->com, 1.0 / M);
    for (int i = 1; i < m->nbody; i++) {
        double D[9] = {m->inertia[i][0], 0, 0, 0, m->inertia[i][1], 0, 0, 0, m->inertia[i][2]}, T[9], Rt[9];
        m3_transpose(Rt, d->ximat[i]); m3_mul(T, d->ximat[i], D); m3_mul(T, T, Rt);
        double r[3]; v3_sub(r, d->xipos[i], d->com);
        double mm = m->mass[i], r2 = v3_dot(r, r);
        double *c = d->cinert[i];
        c[0] = T[0] + mm * (r2 - r[0] * r[0]); c[1] = T[4] + mm * (r2 - r[1] * r[1]); c[2] = T[8] + mm * (r2 - r[2] * r[2]);
        c[3] = T[1] - mm * r[0] * r[1]; c[4] = T[2] - mm * r[0] * r[2]; c[5] = T[5] - mm * r[1] * r[2];
        c[6] = mm * r[0]; c[7] = mm * r[1]; c[8] = mm * r[2]; c[9] = mm;
    }
    /* cdof (mj_dofCom): free joint -> 3 world-axis translations, 3 body-axis rotations */
    for (int k = 0; k < 3; k++) d->cdof[k][3 + k] = 1.0;
    for (int k = 0; k < 3; k++) {
        double axis[3] = {d->xmat[1][k], d->xmat[1][3 + k], d->xmat[1][6 + k]}, off[3], l[3];
        v3_sub(off, d->com, d->xanchor[1]); v3_cross(l, axis, off);
        memcpy(d->cdof[3 + k], axis, sizeof axis); memcpy(d->cdof[3 + k] + 3, l, sizeof l);
    }
    for (int i = 3; i < m->nbody; i++) {
        double off[3], l[3];
        v3_sub(off, d->com, d->xanchor[i]); v3_cross(l, d->xaxis[i], off);
        memcpy(d->cdof[m->dofadr[i]], d->xaxis[i], 3 * sizeof(double)); memcpy(d->cdof[m->dofadr[i]] + 3, l, sizeof l);
    }
}

/* mj_crb: composite inertias, dense symmetric M */
static void crb(const OrcModel *m, OData *d) {
    double c[ORC_MAXBODY][10];
    memcpy(c, d->cinert, sizeof c);
    for (int i = m->nbody - 1; i > 1; i--) for (int k = 0; k < 10; k++) c[m->parent[i]][k] += c[i][k];
    int nv = m->nv;
    for (int i = 0; i < nv; i++) {
        double buf[6]; inert_mulvec(buf, c[m->dof_body[i]], d->cdof[i]);
        for (int j = 0; j <= i; j++) {
            /* j must be an ancestor dof of i: true for all j<=i in this serial chain */
            double v = dot6(d->cdof[j], buf);
            d->M[i * nv + j] = v; d->M[j * nv + i] = v;
        }
    }
}

/* mj_comVel */
static void com_vel(const OrcModel *m, const double *qvel, OData *d) {
    for (int i = 1; i < m->nbody; i++) {
        double cvel[6]; memcpy(cvel, d->cvel[m->parent[i]], sizeof cvel);
        if (m->jnt_type[i] == 1) {
            for (int k = 0; k < 3; k++) for (int r = 0; r < 6; r++) cvel[r] += d->cdof[k][r] * qvel[k];   /* translations: cdof_dot = 0 */
            for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot[k], cvel, d->cdof[k]);
            for (int k = 3; k < 6; k++) for (int r = 0; r < 6; r++) cvel[r] += d->cdof[k][r] * qvel[k];
        } else if (m->jnt_type[i] == 2) {
            int a = m->dofadr[i];
            cross_motion(d->cdof_dot[a], cvel, d->cdof[a]);
            for (int r = 0; r < 6; r++) cvel[r] += d->cdof[a][r] * qvel[a];
        }
        memcpy(d->cvel[i], cvel, sizeof cvel);
    }
}

/* mj_rne: flg_acc=0 -> bias; with qacc!=NULL also returns per-body cacc (mj_rnePostConstraint part) */
static void rne(const OrcModel *m, const double *qvel, const double *qacc, OData *d, double *result, double cacc_out[][6]) {
    double cacc[ORC_MAXBODY][6], cfrc[ORC_MAXBODY][6];
    memset(cacc, 0, sizeof cacc); memset(cfrc, 0, sizeof cfrc);
    cacc[0][3] = -m->gravity[0]; cacc[0][4] = -m->gravity[1]; cacc[0][5] = -m->gravity[2];
    for (int i = 1; i < m->nbody; i++) {
        memcpy(cacc[i], cacc[m->parent[i]], sizeof cacc[i]);
        int a = m->dofadr[i], n = m->jnt_type[i] == 1 ? 6 : (m->jnt_type[i] == 2 ? 1 : 0);
        for (int k = 0; k < n; k++)
            for (int r = 0; r < 6; r++) {
                cacc[i][r] += d->cdof_dot[a + k][r] * qvel[a + k];
                if (qacc) cacc[i][r] += d->cdof[a + k][r] * qacc[a + k];
            }
        double t1[6], t2[6], t3[6];
        inert_mulvec(t1, d->cinert[i], cacc[i]);
        inert_mulvec(t2, d->cinert[i], d->cvel[i]);
        cross_force(t3, d->cvel[i], t2);
        for (int r = 0; r < 6; r++) cfrc[i][r] = t1[r] + t3[r];
    }
    for (int i = m->nbody - 1; i > 1; i--) for (int r = 0; r < 6; r++) cfrc[m->parent[i]][r] += cfrc[i][r];
    if (result) for (int k = 0; k < m->nv; k++) result[k] = dot6(d->cdof[k], cfrc[m->dof_body[k]]);
    if (cacc_out) memcpy(cacc_out, cacc, sizeof cacc);
}

/* mj_jac-style projection: qfrc += J_p^T f + J_r^T t for a point on `body` */
static void apply_ft(const OrcModel *m, const OData *d, const double *f, const double *t, const double *point, int body, double *qfrc) {
    double off[3]; v3_sub(off, point, d->com);
    for (int k = 0; k < m->nv; k++) {
        /* dof k affects `body` iff its body is an ancestor-or-self: bodies are a serial chain 1<2<3<4 */
        if (m->dof_body[k] > body) continue;
        double jp[3], tmp[3];
        v3_cross(tmp, d->cdof[k], off); v3_add(jp, d->cdof[k] + 3, tmp);
        qfrc[k] += v3_dot(jp, f) + v3_dot(d->cdof[k], t);
    }
}

/* mj_passive: joint damping + mj_inertiaBoxFluidModel for every body with mass */
static void passive(const OrcModel *m, const double *qvel, OData *d) {
    for (int k = 0; k < m->nv; k++) d->qfrc_passive[k] = -m->damping[k] * qvel[k];
    if (m->density <= 0 && m->viscosity <= 0) return;
    for (int i = 1; i < m->nbody; i++) {
        if (m->mass[i] < MJMINVAL) continue;
        const double *I = m->inertia[i];
        double box[3];
        box[0] = sqrt(fmax(MJMINVAL, I[1] + I[2] - I[0]) / m->mass[i] * 6.0);
        box[1] = sqrt(fmax(MJMINVAL, I[0] + I[2] - I[1]) / m->mass[i] * 6.0);
        box[2] = sqrt(fmax(MJMINVAL, I[0] + I[1] - I[2]) / m->mass[i] * 6.0);
        /* mj_objectVelocity(mjOBJ_BODY, flg_local=1): 6D velocity at xipos in the inertial frame */
        double off[3], wxr[3], vlin[3], lvel[6];
        v3_sub(off, d->xipos[i], d->com); v3_cross(wxr, d->cvel[i], off); v3_add(vlin, d->cvel[i] + 3, wxr);
        m3_tmulv(lvel, d->ximat[i], d->cvel[i]); m3_tmulv(lvel + 3, d->ximat[i], vlin);
        double lfrc[6] = {0};
        if (m->viscosity > 0) {
            double diam = (box[0] + box[1] + box[2]) / 3.0;
            for (int k = 0; k < 3; k++) lfrc[k] = -PI * diam * diam * diam * m->viscosity * lvel[k];
            for (int k = 0; k < 3; k++) lfrc[3 + k] = -3.0 * PI * diam * m->viscosity * lvel[3 + k];
        }
        if (m->density > 0) {
            lfrc[3] -= 0.5 * m->density * box[1] * box[2] * fabs(lvel[3]) * lvel[3];
            lfrc[4] -= 0.5 * m->density * box[0] * box[2] * fabs(lvel[4]) * lvel[4];
            lfrc[5] -= 0.5 * m->density * box[0] * box[1] * fabs(lvel[5]) * lvel[5];
            lfrc[0] -= m->density * box[0] * (pow(box[1], 4) + pow(box[2], 4)) * fabs(lvel[0]) * lvel[0] / 64.0;
            lfrc[1] -= m->density * box[1] * (pow(box[0], 4) + pow(box[2], 4)) * fabs(lvel[1]) * lvel[1] / 64.0;
            lfrc[2] -= m->density * box[2] * (pow(box[0], 4) + pow(box[1], 4)) * fabs(lvel[2]) * lvel[2] / 64.0;
        }
        double bt[3], bf[3];
        m3_mulv(bt, d->ximat[i], lfrc); m3_mulv(bf, d->ximat[i], lfrc + 3);
        apply_ft(m, d, bf, bt, d->xipos[i], i, d->qfrc_passive);
    }
}

static void solve_dense(int n, const double *A, const double *b, double *x) {
    double a[ORC_MAXNV * ORC_MAXNV], r[ORC_MAXNV];
    memcpy(a, A, n * n * sizeof(double)); memcpy(r, b, n * sizeof(double));
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int i = c + 1; i < n; i++) if (fabs(a[i * n + c]) > fabs(a[piv * n + c])) piv = i;
        if (piv != c) { for (int j = 0; j < n; j++) { double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t; } double t = r[c]; r[c] = r[piv]; r[piv] = t; }
        for (int i = c + 1; i < n; i++) {
            double f = a[i * n + c] / a[c * n + c];
            for (int j = c; j < n; j++) a[i * n + j] -= f * a[c * n + j];
            r[i] -= f * r[c];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = r[i];
        for (int j = i + 1; j < n; j++) s -= a[i * n + j] * x[j];
        x[i] = s / a[i * n + i];
    }
}


/* ------------------------------------------------------------------ floor contact (models with `ground` set)
 * env_gen.py:14-21: every drone geom has contype 1 / conaffinity 0, condim 3, friction (1, .5, .5), margin 0; the floor
 * (:97) is a default plane geom (contype = conaffinity = 1, friction (1, .005, .0001)) at z = 0.  Drone geoms therefore
 * collide with the floor and with nothing else.  Contact parameters are MuJoCo's defaults on both sides: solref (0.02, 1),
 * solimp (0.9, 0.95, 0.001, 0.5, 2), friction = element-wise max -> mu = 1, condim 3, pyramidal cone, impratio 1.
 * [MuJoCo's algorithm restated from its documentation / engine_collision_primitive.c / engine_core_constraint.c; unpinned.] */
static void geom_pose(const OData *d, const OrcGeom *g, double gp[3], double gm[9]) {
    double t[3]; m3_mulv(t, d->xmat[g->body], g->pos); v3_add(gp, d->xpos[g->body], t);
    double c = cos(g->yaw), s = sin(g->yaw), Rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    m3_mul(gm, d->xmat[g->body], Rz);
}
static int add_con(OrcContact *con, int n, const double p[3], double dist, int body, int geom) {
    if (n >= ORC_MAXCON) return n;
    /* contact position: half-way between the geom's point and the plane surface */
    con[n].pos[0] = p[0]; con[n].pos[1] = p[1]; con[n].pos[2] = p[2] - 0.5 * dist;
    con[n].dist = dist; con[n].body = body; con[n].geom = geom;
    return n + 1;
}
static int collide_floor(const OrcModel *m, const OData *d, OrcContact *con) {
    const double margin = 0.0;
    int n = 0;
    for (int gi = 0; gi < m->ngeom; gi++) {
        const OrcGeom *g = &m->geom[gi];
        double gp[3], gm[9];
        geom_pose(d, g, gp, gm);
        const double dist0 = gp[2];                                   /* plane through the origin, normal +z */
        if (g->type == 2) {                                           /* mjc_PlaneSphere */
            double dist = dist0 - g->size[0];
            if (dist > margin) continue;
            double p[3] = {gp[0], gp[1], gp[2] - g->size[0]};
            n = add_con(con, n, p, dist, g->body, gi);
        } else if (g->type == 0) {                                    /* mjc_PlaneBox: corners below the plane that point down, at most 4 */
            int cnt = 0;
            for (int i = 0; i < 8 && cnt < 4; i++) {
                double v[3] = {(i & 1 ? 1 : -1) * g->size[0], (i & 2 ? 1 : -1) * g->size[1], (i & 4 ? 1 : -1) * g->size[2]}, c[3];
                m3_mulv(c, gm, v);
                double ldist = c[2];
                if (dist0 + ldist > margin || ldist > 0) continue;
                double p[3] = {gp[0] + c[0], gp[1] + c[1], gp[2] + c[2]};
                n = add_con(con, n, p, dist0 + ldist, g->body, gi); cnt++;
            }
        } else {                                                      /* mjc_PlaneCylinder */
            double axis[3] = {gm[2], gm[5], gm[8]};
            double prjaxis = axis[2];
            if (prjaxis > 0) { v3_scl(axis, axis, -1.0); prjaxis = -prjaxis; }
            /* direction inside the disc plane that points most steeply towards the floor: -z projected */
            double vec[3] = {axis[0] * prjaxis, axis[1] * prjaxis, axis[2] * prjaxis - 1.0};
            double len = sqrt(v3_dot(vec, vec));
            if (len < 1e-12) { vec[0] = gm[0] * g->size[0]; vec[1] = gm[3] * g->size[0]; vec[2] = gm[6] * g->size[0]; }   /* disc parallel to the floor */
            else v3_scl(vec, vec, g->size[0] / len);
            double prjvec = vec[2];
            v3_scl(axis, axis, g->size[1]); prjaxis *= g->size[1];
            if (dist0 + prjaxis + prjvec > margin) continue;
            double p[3];
            for (int k = 0; k < 3; k++) p[k] = gp[k] + axis[k] + vec[k];
            n = add_con(con, n, p, dist0 + prjaxis + prjvec, g->body, gi);
            if (dist0 - prjaxis + prjvec <= margin) {                 /* same direction on the far cap */
                for (int k = 0; k < 3; k++) p[k] = gp[k] - axis[k] + vec[k];
                n = add_con(con, n, p, dist0 - prjaxis + prjvec, g->body, gi);
            }
            double prjvec1 = -0.5 * prjvec;                           /* two more points of the near rim, 120 degrees away */
            if (dist0 + prjaxis + prjvec1 <= margin) {
                double side[3]; v3_cross(side, vec, axis);
                double sl = sqrt(v3_dot(side, side));
                v3_scl(side, side, g->size[0] * sqrt(3.0) / 2 / (sl > 1e-300 ? sl : 1e-300));
                for (int sg = -1; sg <= 1; sg += 2) {
                    for (int k = 0; k < 3; k++) p[k] = gp[k] + axis[k] - 0.5 * vec[k] + sg * side[k];
                    n = add_con(con, n, p, dist0 + prjaxis + prjvec1, g->body, gi);
                }
            }
        }
    }
    return n;
}

/* translational Jacobian of world point `point` moving with `body`: jp[k] for every dof k (mj_jac) */
static void jac_point(const OrcModel *m, const OData *d, const double *point, int body, double jp[][3]) {
    double off[3]; v3_sub(off, point, d->com);
    for (int k = 0; k < m->nv; k++) {
        if (m->dof_body[k] > body) { v3_zero(jp[k]); continue; }
        double tmp[3]; v3_cross(tmp, d->cdof[k], off); v3_add(jp[k], d->cdof[k] + 3, tmp);
    }
}

/* mjModel.body_invweight0 (engine_setconst.c, set0): at qpos0, A = J M^-1 J^T with the 6 x nv Jacobian of the body at its
 * COM; translation weight = mean of the first three diagonal entries, rotation weight = mean of the last three */
static void set_invweight0(OrcModel *m) {
    double qpos0[ORC_MAXNQ] = {0, 0, 0, 1, 0, 0, 0, 0, 0};
    OData d;
    kin_com(m, qpos0, &d); crb(m, &d);
    int nv = m->nv;
    for (int b = 2; b < m->nbody; b++) {
        double jp[ORC_MAXNV][3], tr = 0, rot = 0;
        jac_point(m, &d, d.xipos[b], b, jp);
        for (int a = 0; a < 3; a++) {
            double col[ORC_MAXNV], x[ORC_MAXNV];
            for (int k = 0; k < nv; k++) col[k] = jp[k][a];
            solve_dense(nv, d.M, col, x);
            for (int k = 0; k < nv; k++) tr += col[k] * x[k];
            for (int k = 0; k < nv; k++) col[k] = m->dof_body[k] > b ? 0.0 : d.cdof[k][a];
            solve_dense(nv, d.M, col, x);
            for (int k = 0; k < nv; k++) rot += col[k] * x[k];
        }
        m->invweight0[b][0] = fmax(MJMINVAL, tr / 3); m->invweight0[b][1] = fmax(MJMINVAL, rot / 3);
    }
}

/* mj_makeConstraint + mj_makeImpedance for the contacts, then the convex problem of mj_fwdConstraint:
 *   qacc = argmin_x  1/2 (x - qacc_smooth)^T M (x - qacc_smooth) + sum_rows 1/2 D_r min(0, J_r x - aref_r)^2
 * (primal form, pyramidal cone: every row is a one-sided soft constraint).  The minimiser is unique (M > 0), so any solver
 * that converges reproduces MuJoCo's Newton solver up to its tolerance; this one is Newton with an exact line search. */
static void fwd_constraint(const OrcModel *m, const double *qvel, OData *d, double *qacc) {
    static const double solref[2] = {0.02, 1.0}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2.0};
    const int nv = m->nv;
    OrcContact con[ORC_MAXCON];
    memset(d->qfrc_constraint, 0, sizeof d->qfrc_constraint);
    const int ncon = collide_floor(m, d, con);
    d->ncon = ncon;
    if (!ncon) return;
    const int nrow = 4 * ncon;
    static __thread double J[4 * ORC_MAXCON][ORC_MAXNV], aref[4 * ORC_MAXCON], Dv[4 * ORC_MAXCON], jar[4 * ORC_MAXCON], jd[4 * ORC_MAXCON];
    /* reference acceleration parameters (mj_makeImpedance: getsolparam), refsafe: timeconst >= 2 timestep */
    const double tc = fmax(solref[0], 2 * m->timestep), dr = solref[1], dmax = solimp[1];
    const double K = 1.0 / fmax(MJMINVAL, dmax * dmax * tc * tc * dr * dr), B = 2.0 / fmax(MJMINVAL, dmax * tc);
    const double mu = 1.0;                                           /* max(1, 1) / sqrt(impratio = 1) */
    for (int c = 0; c < ncon; c++) {
        double jp[ORC_MAXNV][3];
        jac_point(m, d, con[c].pos, con[c].body, jp);
        /* impedance d(r): smooth step of |r| / width between solimp[0] and solimp[1] (midpoint, power) */
        double x = fabs(con[c].dist) / solimp[2], y;
        if (x >= 1) y = 1;
        else if (x <= solimp[3]) y = pow(x, solimp[4]) / pow(solimp[3], solimp[4] - 1);
        else y = 1 - pow(1 - x, solimp[4]) / pow(1 - solimp[3], solimp[4] - 1);
        const double imp = solimp[0] + y * (solimp[1] - solimp[0]);
        /* mj_diagApprox, pyramidal rows: tran + mu^2 tran with tran = invweight0 of the two bodies (the world's is 0);
         * R = (1 - imp) / imp * diagApprox, then all rows of the contact get Rpy = 2 mu^2 R */
        const double tran = m->invweight0[con[c].body][0];
        const double R = 2 * mu * mu * fmax(MJMINVAL, (1 - imp) / imp * (tran + mu * mu * tran));
        /* contact frame of a +z normal (mju_makeFrame): t1 = +y, t2 = -x; rows n + mu t1, n - mu t1, n + mu t2, n - mu t2 */
        static const double dirs[4][3] = {{0, 1, 1}, {0, -1, 1}, {-1, 0, 1}, {1, 0, 1}};
        for (int r = 0; r < 4; r++) {
            const int i = 4 * c + r;
            double vel = 0;
            for (int k = 0; k < nv; k++) { J[i][k] = mu * (dirs[r][0] * jp[k][0] + dirs[r][1] * jp[k][1]) + jp[k][2]; vel += J[i][k] * qvel[k]; }
            Dv[i] = 1.0 / R;
            aref[i] = -B * vel - K * imp * con[c].dist;
        }
    }
    /* Newton */
    double x[ORC_MAXNV], g[ORC_MAXNV], H[ORC_MAXNV * ORC_MAXNV], dx[ORC_MAXNV], Mx[ORC_MAXNV] = {0};
    memcpy(x, qacc, nv * sizeof(double));
    double scale = 0;
    for (int k = 0; k < nv; k++) scale += d->M[k * nv + k];
    scale = 1.0 / (scale / nv * nv);                                  /* 1 / (meaninertia * nv), as MuJoCo scales its tolerance */
    for (int it = 0; it < 100; it++) {
        for (int k = 0; k < nv; k++) { Mx[k] = -d->qfrc_smooth[k]; for (int j = 0; j < nv; j++) Mx[k] += d->M[k * nv + j] * x[j]; }
        memcpy(g, Mx, nv * sizeof(double)); memcpy(H, d->M, nv * nv * sizeof(double));
        for (int i = 0; i < nrow; i++) {
            double v = -aref[i];
            for (int k = 0; k < nv; k++) v += J[i][k] * x[k];
            jar[i] = v;
            if (v >= 0) continue;
            for (int k = 0; k < nv; k++) {
                g[k] += Dv[i] * v * J[i][k];
                for (int j = 0; j < nv; j++) H[k * nv + j] += Dv[i] * J[i][k] * J[i][j];
            }
        }
        double gn = 0;
        for (int k = 0; k < nv; k++) gn += g[k] * g[k];
        if (sqrt(gn) * scale < 1e-13) break;
        for (int k = 0; k < nv; k++) g[k] = -g[k];
        solve_dense(nv, H, g, dx);
        /* exact line search on phi'(a) = p0 + a p2 + sum_r D_r jd_r min(0, jar_r + a jd_r): increasing, piecewise linear */
        double p0 = 0, p2 = 0;
        for (int k = 0; k < nv; k++) { p0 += Mx[k] * dx[k]; for (int j = 0; j < nv; j++) p2 += dx[k] * d->M[k * nv + j] * dx[j]; }
        for (int i = 0; i < nrow; i++) { double v = 0; for (int k = 0; k < nv; k++) v += J[i][k] * dx[k]; jd[i] = v; }
        double a = 0;
        for (int ls = 0; ls < 64; ls++) {
            double f1 = p0 + a * p2, f2 = p2, nxt = 1e300;            /* value, slope, next breakpoint beyond a */
            for (int i = 0; i < nrow; i++) {
                /* row i is active (J x - aref < 0) just to the right of a: decided from its breakpoint, not from a rounded residual */
                int act;
                if (jd[i] != 0) {
                    double bp = -jar[i] / jd[i];
                    act = jd[i] < 0 ? (a >= bp) : (a < bp);
                    if (bp > a && bp < nxt) nxt = bp;
                } else act = jar[i] < 0;
                if (act) { double v = jar[i] + a * jd[i]; f1 += Dv[i] * jd[i] * v; f2 += Dv[i] * jd[i] * jd[i]; }
            }
            if (f1 >= 0) break;                                       /* (only by rounding: phi' is negative left of the root) */
            double root = a - f1 / f2;
            if (root <= nxt) { a = root; break; }
            a = nxt;
        }
        double big = 0;
        for (int k = 0; k < nv; k++) { x[k] += a * dx[k]; big = fmax(big, fabs(a * dx[k]) / (1 + fabs(x[k]))); }
        if (big < 1e-15) break;
    }
    memcpy(qacc, x, nv * sizeof(double));
    for (int k = 0; k < nv; k++) {
        double f = 0;
        for (int i = 0; i < nrow; i++) {
            double v = -aref[i];
            for (int j = 0; j < nv; j++) v += J[i][j] * x[j];
            if (v < 0) f -= Dv[i] * v * J[i][k];
        }
        d->qfrc_constraint[k] = f;
    }
}

int orc_collide(const OrcModel *m, const double *qpos, OrcContact *con) {
    OData d;
    kin_com(m, qpos, &d);
    return collide_floor(m, &d, con);
}

static void forward_core(const OrcModel *m, const double *qpos, const double *qvel, const double *act,
                         const double *ctrl, OData *d, double *qacc, double *act_dot, double *sensordata) {
    int nv = m->nv;
    kin_com(m, qpos, d);
    crb(m, d);
    com_vel(m, qvel, d);
    passive(m, qvel, d);
    rne(m, qvel, NULL, d, d->qfrc_bias, NULL);
    /* mj_fwdActuation: ctrl clamped to ctrlrange (0,1); dyntype=filter; gain 1, no bias -> force = act */
    memset(d->qfrc_actuator, 0, sizeof d->qfrc_actuator);
    for (int k = 0; k < 4; k++) {
        double c = ctrl[k] < 0 ? 0 : (ctrl[k] > 1 ? 1 : ctrl[k]);
        act_dot[k] = (c - act[k]) / fmax(MJMINVAL, m->tau[k]);
        double force = act[k];
        double fw[3], tw[3];
        m3_mulv(fw, d->xmat[2], m->gear[k]); m3_mulv(tw, d->xmat[2], m->gear[k] + 3);
        v3_scl(fw, fw, force); v3_scl(tw, tw, force);
        apply_ft(m, d, fw, tw, d->site_xpos[k], 2, d->qfrc_actuator);
    }
    for (int k = 0; k < nv; k++) d->qfrc_smooth[k] = d->qfrc_passive[k] - d->qfrc_bias[k] + d->qfrc_actuator[k];
    solve_dense(nv, d->M, d->qfrc_smooth, qacc);       /* qacc_smooth; stays the answer when no contact is active */
    if (m->ground) fwd_constraint(m, qvel, d, qacc);
    /* accelerometer (mj_sensorAcc): site 'sense' on body 2 */
    double cacc[ORC_MAXBODY][6];
    rne(m, qvel, qacc, d, NULL, cacc);
    const double *sp = d->site_xpos[4];
    double off[3], a_ang[3], a_lin[3], v_lin[3], t[3];
    v3_sub(off, sp, d->com);
    v3_copy(a_ang, cacc[2]); v3_cross(t, cacc[2], off); v3_add(a_lin, cacc[2] + 3, t);
    v3_cross(t, d->cvel[2], off); v3_add(v_lin, d->cvel[2] + 3, t);
    v3_cross(t, d->cvel[2], v_lin); v3_addto(a_lin, t);          /* acc_tran += omega x v */
    (void)a_ang;
    m3_tmulv(sensordata, d->xmat[2], a_lin);
}

void orc_forward(const OrcModel *m, const double *qpos, const double *qvel, const double *act,
                 const double *ctrl, double *qacc, double *act_dot, double *sensordata,
                 double *M_out, double *qfrc_smooth_out) {
    OData d;
    forward_core(m, qpos, qvel, act, ctrl, &d, qacc, act_dot, sensordata);
    if (M_out) memcpy(M_out, d.M, m->nv * m->nv * sizeof(double));
    if (qfrc_smooth_out) memcpy(qfrc_smooth_out, d.qfrc_smooth, m->nv * sizeof(double));
}

int orc_forward_contact(const OrcModel *m, const double *qpos, const double *qvel, const double *act, const double *ctrl,
                        double *qacc, double *qfrc_constraint, double *sensordata) {
    OData d;
    double act_dot[4];
    forward_core(m, qpos, qvel, act, ctrl, &d, qacc, act_dot, sensordata);
    if (qfrc_constraint) memcpy(qfrc_constraint, d.qfrc_constraint, m->nv * sizeof(double));
    return m->ground ? d.ncon : 0;
}

void orc_step(const OrcModel *m, double *qpos, double *qvel, double *act, const double *ctrl,
              double *sensordata, int nstep) {
    int nv = m->nv;
    double h = m->timestep;
    for (int s = 0; s < nstep; s++) {
        OData d;
        double qacc[ORC_MAXNV], act_dot[4], qacc_i[ORC_MAXNV];
        forward_core(m, qpos, qvel, act, ctrl, &d, qacc, act_dot, sensordata);
        /* mj_EulerSkip: implicit in joint damping when any dof_damping > 0 */
        int damp = 0;
        for (int k = 0; k < nv; k++) if (m->damping[k] > 0) damp = 1;
        if (damp) {
            double H[ORC_MAXNV * ORC_MAXNV];
            memcpy(H, d.M, nv * nv * sizeof(double));
            for (int k = 0; k < nv; k++) H[k * nv + k] += h * m->damping[k];
            double f[ORC_MAXNV];
            for (int k = 0; k < nv; k++) f[k] = d.qfrc_smooth[k] + d.qfrc_constraint[k];
            solve_dense(nv, H, f, qacc_i);
        } else memcpy(qacc_i, qacc, nv * sizeof(double));
        /* mj_advance */
        for (int k = 0; k < 4; k++) act[k] += h * act_dot[k];
        for (int k = 0; k < nv; k++) qvel[k] += h * qacc_i[k];
        for (int k = 0; k < 3; k++) qpos[k] += h * qvel[k];
        double ax[3] = {qvel[3], qvel[4], qvel[5]};
        double n = sqrt(v3_dot(ax, ax));
        if (n < MJMINVAL) { ax[0] = 1; ax[1] = 0; ax[2] = 0; } else { ax[0] /= n; ax[1] /= n; ax[2] /= n; }
        double qr[4]; q_axis_angle(qr, ax, h * n);
        q_normalize(qpos + 3);
        q_mul(qpos + 3, qpos + 3, qr);
        for (int k = 6; k < nv; k++) qpos[k + 1] += h * qvel[k];
    }
}

void orc_energy(const OrcModel *m, const double *qpos, const double *qvel, double *ke, double *pe) {
    OData d;
    kin_com(m, qpos, &d); crb(m, &d);
    double k = 0;
    for (int i = 0; i < m->nv; i++) for (int j = 0; j < m->nv; j++) k += 0.5 * qvel[i] * d.M[i * m->nv + j] * qvel[j];
    double p = 0;
    for (int i = 1; i < m->nbody; i++) p -= m->mass[i] * v3_dot(m->gravity, d.xipos[i]);
    *ke = k; *pe = p;
}

/* ------------------------------------------------------------------ transformation.py */
static double pymod(double a, double b) { double r = fmod(a, b); if (r != 0 && ((r < 0) != (b < 0))) r += b; return r; }
static double wrap_pi(double a) { return pymod(a + PI, 2 * PI) - PI; }

/* transformation.py:15-17: R.from_quat([x,y,z,w]).as_euler('ZYX')[::-1]; scipy's quaternion algorithm
 * (intrinsic ZYX -> i=0,j=1,k=2, sign=+1, not symmetric), incl. gimbal-lock branches, eps = 1e-7 */
void orc_quat2rpy(const double qw[4], double rpy[3]) {
    double q[4] = {qw[0], qw[1], qw[2], qw[3]};
    q_normalize(q);
    double w = q[0], x = q[1], y = q[2], z = q[3];
    double a = w - y, b = x + z, c = y + w, d = z - x;
    double half_sum = atan2(b, a), half_diff = atan2(d, c);
    double ang1 = 2 * atan2(hypot(c, d), hypot(a, b));
    int case1 = fabs(ang1) <= 1e-7, case2 = fabs(ang1 - PI) <= 1e-7;
    double ang0, ang2;  /* ang0 = first of seq (yaw, Z), ang2 = last (roll, X) */
    if (!case1 && !case2) { ang2 = half_sum - half_diff; ang0 = half_sum + half_diff; }
    else { ang2 = 0.0; ang0 = case1 ? 2 * half_sum : 2 * half_diff; }
    ang1 -= PI / 2;
    rpy[0] = wrap_pi(ang2); rpy[1] = wrap_pi(ang1); rpy[2] = wrap_pi(ang0);
}
/* transformation.py:20-23: R.from_euler('ZYX',[yaw,pitch,roll]) = qz(yaw) * qy(pitch) * qx(roll) */
void orc_rpy2quat(const double rpy[3], double q[4]) {
    double cr = cos(rpy[0] / 2), sr = sin(rpy[0] / 2), cp = cos(rpy[1] / 2), sp = sin(rpy[1] / 2), cy = cos(rpy[2] / 2), sy = sin(rpy[2] / 2);
    double qz[4] = {cy, 0, 0, sy}, qy[4] = {cp, 0, sp, 0}, qx[4] = {cr, sr, 0, 0}, t[4];
    q_mul(t, qz, qy); q_mul(q, t, qx);
}
void orc_quat2dcm(const double qw[4], double R[9]) { double q[4] = {qw[0], qw[1], qw[2], qw[3]}; q_normalize(q); q_to_mat(R, q); }
/* transformation.py:26-28: from_euler('XY') = qx(r) * qy(p) */
void orc_pendulumrp2quat(const double rp[2], double q[4]) {
    double qx[4] = {cos(rp[0] / 2), sin(rp[0] / 2), 0, 0}, qy[4] = {cos(rp[1] / 2), 0, sin(rp[1] / 2), 0};
    q_mul(q, qx, qy);
}
static void rpy2dcm(const double rpy[3], double R[9]) { double q[4]; orc_rpy2quat(rpy, q); orc_quat2dcm(q, R); }

/* ------------------------------------------------------------------ BaseDroneEnv.get_drone_states (:357-380) */
int orc_drone_state(const OrcModel *m, const double *qpos, const double *qvel, const double *act,
                    const double *sens, const double ref[4], double *s) {
    int n = 0;
    for (int k = 0; k < 3; k++) s[n++] = qpos[k];
    double rpy[3]; orc_quat2rpy(qpos + 3, rpy);
    for (int k = 0; k < 3; k++) s[n++] = rpy[k];
    for (int k = 0; k < 6; k++) s[n++] = qvel[k];
    if (m->nq == 9) { s[n++] = qpos[7]; s[n++] = qpos[8]; s[n++] = qvel[6]; s[n++] = qvel[7]; }
    for (int k = 0; k < 3; k++) s[n++] = sens[k];
    for (int k = 0; k < 4; k++) s[n++] = act[k];
    for (int k = 0; k < 4; k++) s[n++] = ref[k];
    for (int k = 0; k < 6; k++) s[n++] = m->params[k];
    return n;
}

/* BaseDroneEnv.py:12-16.  Sum of squares in index order, no FMA contraction (compile with -ffp-contract=off) */
int orc_termination(const double *s, const double ref[4], double max_distance, int64_t num_steps, int64_t max_steps) {
    double dx = s[0] - ref[0], dy = s[1] - ref[1], dz = s[2] - ref[2];
    double pos_err = sqrt((dx * dx + dy * dy) + dz * dz);
    return (pos_err > max_distance) || (num_steps >= max_steps);
}

/* ------------------------------------------------------------------ rewards.py */
static double sq(double x) { return x * x; }
static double heading_wrapped(const double *s, const double *ref) { return pymod(fabs(s[5] - ref[3]) + PI, 2 * PI) - PI; }
static double pos_err_sq(const double *s, const double *ref) { return sq(s[0] - ref[0]) + sq(s[1] - ref[1]) + sq(s[2] - ref[2]); }
static void rot_x(double a, double R[9]) { double c = cos(a), s = sin(a); double t[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(R, t, sizeof t); }
static void rot_y(double a, double R[9]) { double c = cos(a), s = sin(a); double t[9] = {c, 0, s, 0, 1, 0, -s, 0, c}; memcpy(R, t, sizeof t); }

/* rewards.py:81-103: pendulum tip velocity in the world frame (the `_en*` family) */
static void pend_tip_en(const double *s, double pvg[3], double *p_h) {
    const double *params = s + 27, *p_rp = s + 12, *rpy = s + 3, *omega_rp = s + 14, *omega = s + 9;
    double Rd[9], Rp[9], Rx[9], Ry[9], q[4];
    rpy2dcm(rpy, Rd);
    orc_pendulumrp2quat(p_rp, q); orc_quat2dcm(q, Rp);
    rot_x(p_rp[0], Rx); rot_y(p_rp[1], Ry);
    double pe[3] = {0, 0, -params[4]};
    double Ox[9] = {0, 0, 0, 0, 0, -omega_rp[0], 0, omega_rp[0], 0};
    double Oy[9] = {0, 0, omega_rp[1], 0, 0, 0, -omega_rp[1], 0, 0};
    double Oc[9] = {0, -omega[2], omega[1], omega[2], 0, -omega[0], -omega[1], omega[0], 0};
    double T1[9], T2[9], T3[9], v1[3], v2[3];
    m3_mul(T1, Rd, Oc); m3_mul(T1, T1, Rp); m3_mulv(v1, T1, pe);
    m3_mul(T2, Rx, Ox); m3_mul(T2, T2, Ry);
    m3_mul(T3, Rx, Ry); m3_mul(T3, T3, Oy);
    for (int k = 0; k < 9; k++) T2[k] += T3[k];
    m3_mul(T2, Rd, T2); m3_mulv(v2, T2, pe);
    for (int k = 0; k < 3; k++) pvg[k] = v1[k] + v2[k];       /* WITHOUT state[6:9]: see pend_en_sum */
    if (p_h) { double T[9], r[3]; m3_mul(T, Rd, Rp); m3_mulv(r, T, pe); *p_h = r[2]; }
}
/* rewards.py:102-103: `state[6:9] + (3,1) column` BROADCASTS to a 3x3 matrix M[i][j] = vel[j] + w[i];
 * `(pendulum_v_global**2).sum()` then sums all nine entries.  Replicated as is (reference quirk). */
static double pend_en_sum(const double *s, const double w[3]) {
    double e = 0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) e += sq(s[6 + j] + w[i]);
    return e;
}
/* rewards.py:283-292 etc.: pendulum tip position with pendulum_R = DCM(rpy2quat([r,p,0])) */
static void pend_tip_pos(const double *s, double plen, double pos[3], double Rd[9], double Rpend[9]) {
    double prpy[3] = {s[12], s[13], 0};
    rpy2dcm(s + 3, Rd); rpy2dcm(prpy, Rpend);
    double T[9], pe[3] = {0, 0, -plen}, r[3];
    m3_mul(T, Rd, Rpend); m3_mulv(r, T, pe);
    for (int k = 0; k < 3; k++) pos[k] = s[k] + r[k];
}
static double pend_energy_local(const double *s, const double Rd[9], const double Rpend[9]) {
    /* rewards.py:337-341 */
    double pe[3] = {0, 0, -s[27 + 4]}, r[3], om[3] = {s[14], s[15], 0}, vl[3], vg[3];
    m3_mulv(r, Rpend, pe); v3_cross(vl, om, r); m3_mulv(vg, Rd, vl);
    return sq(s[6] + vg[0]) + sq(s[7] + vg[1]) + sq(s[8] + vg[2]);
}

double orc_reward(int id, const double *s, int nstate, const double a[4], int64_t num_steps, const double ref[4], double max_distance) {
    (void)nstate;
    double he = heading_wrapped(s, ref);
    double pe2 = pos_err_sq(s, ref), pe = sqrt(pe2);
    double a2 = sq(a[0]) + sq(a[1]) + sq(a[2]) + sq(a[3]);
    switch (id) {
    case 0: return 3 - pe;                                                                    /* default_reward_fcn :5-10 */
    case 1: return 5 - pe - 0.1 * fabs(he);                                                   /* distance_reward_fcn :13-20 */
    case 2: return 3.5 - pe2 - 0.1 * fabs(he) - 0.2 * a2;                                     /* distance_energy_reward :23-31 */
    case 3: return 3.5 - pe2 - 0.2 * he * he - 0.2 * a2 - 0.2 * (sq(s[12]) + sq(s[13]));      /* ..._pendulum_angle :34-43 */
    case 4: return 3.5 - pe2 - 0.5 * he * he - 0.4 * a2 - 0.2 * (sq(s[12]) + sq(s[13])) - 0.1 * (sq(s[9]) + sq(s[10]) + sq(s[11]));
    case 5: {                                                                                 /* ..._angle3 :59-72 */
        double pd = sq(s[12]) + sq(s[13]), ad = sq(s[3]) + sq(s[4]), rs = sq(s[9]) + sq(s[10]) + sq(s[11]), pav = sq(s[14]) + sq(s[15]);
        double r = 3.5 - pe2 - 0.5 * he * he - 0.4 * a2;
        r -= (0.1 * pd + 0.2 * pav - 0.3 * ad - 0.4 * rs) / (1 + 100 * pe2);
        return r;
    }
    case 6: { double v[3]; pend_tip_en(s, v, NULL); return 3.5 - pe2 - 0.5 * he * he - 0.4 * a2 - 0.2 * pend_en_sum(s, v); }
    case 7: case 8: case 9: {
        double thr = id == 9 ? 0.6 : 0.5, ce = 0;
        for (int k = 0; k < 4; k++) ce += sq(fmax(a[k] - thr, 0));
        double v[3], ph; pend_tip_en(s, v, &ph);
        double en = pend_en_sum(s, v);
        double angdev = sqrt(sq(s[3]) + sq(s[4]) + sq(s[5]));
        if (id == 7) { double r = 3.5 - 2 * pe - 0.6 * he * he - 0.6 * ce; if (pe < 0.15) r = r + 3 - 0.2 * en - 0.2 * angdev; return r; }
        double tot = 0.5 * en + 9.81 * ph;
        if (id == 8) return 7 - pe - 0.4 * he * he - 0.1 * ce - 0.1 * tot - 0.05 * angdev;
        return 5 - pe - 0.6 * he * he - 0.1 * ce - (0.2 * tot + 0.05 * angdev) / (0.5 + pe);
    }
    case 10: {                                                                                /* distance_time_energy_reward :233-242 */
        double too_far = pe2 > max_distance * max_distance ? 1.0 : 0.0;
        return -(1 + (double)(num_steps / 50)) * pe2 - 500 * too_far - fabs(he) - 0.02 * a2;
    }
    case 11: {                                                                                /* reward_1 :245-257 */
        double tilt = sq(s[3]) + sq(s[4]), close = pe2 < 0.2 ? 1.0 : 0.0, rot = sq(s[6]) + sq(s[7]) + sq(s[8]), pen = sq(s[14]) + sq(s[15]);
        double too_far = pe2 > max_distance * max_distance - 3 ? 1.0 : 0.0;
        return (7 + 20 * close - 3 * pe2 * (1 + (double)num_steps / 150) - 10 * too_far - 0.3 * tilt - 0.7 * he * he - 0.3 * a2 - 0.3 * rot - 0.5 * pen) / 10;
    }
    case 12: { double p[3], Rd[9], Rp[9]; pend_tip_pos(s, s[27 + 5], p, Rd, Rp); return -(sq(p[0] - ref[0]) + sq(p[1] - ref[1]) + sq(p[2] - ref[2])); }  /* reward_pendulum_dist :283-294 (params[5], Q12) */
    case 13: { double p[3], Rd[9], Rp[9]; pend_tip_pos(s, s[27 + 4], p, Rd, Rp); double e = sq(p[0] - ref[0]) + sq(p[1] - ref[1]) + sq(p[2] - ref[2]); return 3 - e - 0.1 * fabs(he); }
    case 14: { double p[3], Rd[9], Rp[9]; pend_tip_pos(s, s[27 + 4], p, Rd, Rp); double e = sq(p[0] - ref[0]) + sq(p[1] - ref[1]) + sq(p[2] - ref[2]);
               return 4 - e - 0.001 * (double)num_steps * e - 0.1 * fabs(he) - 0.05 * a2; }                      /* reward_2 :313-327 */
    case 15: { double p[3], Rd[9], Rp[9]; pend_tip_pos(s, s[27 + 4], p, Rd, Rp); double e = sq(p[0] - ref[0]) + sq(p[1] - ref[1]) + sq(p[2] - ref[2]);
               double en = pend_energy_local(s, Rd, Rp), h = fabs(he);
               return 4 - e - 0.2 * h - 0.006 * (double)num_steps * (e + 0.2 * h) - 0.05 * a2 - 0.1 * en; }      /* reward_2_penergy :330-348 */
    case 16: { double p[3], Rd[9], Rp[9]; pend_tip_pos(s, s[27 + 4], p, Rd, Rp);
               double en = pend_energy_local(s, Rd, Rp), h = fabs(he), ce = 0;
               for (int k = 0; k < 4; k++) ce += sq(fmin(a[k] - 0.5, 0));
               return 4 - pe2 - 0.2 * h - 0.006 * (double)num_steps * (pe2 + 0.2 * h + 0.01 * en) - 0.1 * ce - 0.1 * en; } /* reward_3 :351-368 */
    }
    return NAN;
}

/* ------------------------------------------------------------------ observation_wrappers.py */
int orc_obs(int id, const double *s, int nstate, const double ref[4], double *o) {
    if (id == 0) { memcpy(o, s, nstate * sizeof(double)); return nstate; }               /* BaseDroneEnv._get_obs :353-355 */
    if (id == 12) return -1;                                                             /* NameError at observation_wrappers.py:448 (Q14) */
    const double *xyz = s, *rpy = s + 3, *vel = s + 6, *angvel = s + 9, *prp = s + 12, *pav = s + 14, *acc = s + 16, *act = s + 19, *params = s + 27;
    double hd = pymod(ref[3] - rpy[2] + PI, 2 * PI) - PI;
    double gerr[3] = {ref[0] - xyz[0], ref[1] - xyz[1], ref[2] - xyz[2]};
    double R[9], lerr[3], lvel[3];
    rpy2dcm(rpy, R); m3_tmulv(lerr, R, gerr); m3_tmulv(lvel, R, vel);
    double rp0[3] = {rpy[0], rpy[1], 0}, Z[9]; rpy2dcm(rp0, Z);
    double zvec[3] = {Z[2], Z[5], Z[8]};
    int n = 0;
#define PUT3(v) do { o[n++] = (v)[0]; o[n++] = (v)[1]; o[n++] = (v)[2]; } while (0)
#define PUT2(v) do { o[n++] = (v)[0]; o[n++] = (v)[1]; } while (0)
#define PUT2R(v) do { o[n++] = (v)[1]; o[n++] = (v)[0]; } while (0)
    switch (id) {
    case 1: PUT3(gerr); PUT2(rpy); o[n++] = hd; PUT3(vel); PUT3(angvel); PUT2(prp); PUT2(pav); break;                         /* GlobalFrameRPYEnv :7-35 */
    case 2: PUT3(lerr); PUT2R(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT2R(prp); PUT2(pav); break;                      /* LocalFramePRYEnv :38-73 */
    case 3: PUT3(lerr); PUT2R(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT3(acc); for (int k = 0; k < 4; k++) o[n++] = act[k]; PUT2R(prp); PUT2(pav); break; /* FullState :76-111 */
    case 4: PUT3(lerr); PUT3(zvec); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT3(acc); for (int k = 0; k < 4; k++) o[n++] = act[k]; PUT2R(prp); PUT2(pav); break; /* FullStateZvec :114-151 (24 values) */
    case 5: PUT3(lerr); PUT2R(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT3(acc); PUT2R(prp); PUT2(pav); break;           /* PRYacc :154-191 */
    case 6: PUT3(lerr); PUT2R(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT2R(prp); PUT2(pav); for (int k = 0; k < 6; k++) o[n++] = params[k]; break; /* PRYParams :194-230 */
    case 7: PUT3(lerr); PUT2R(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT2R(prp); PUT3(acc); PUT2(pav); for (int k = 0; k < 6; k++) o[n++] = params[k]; break; /* PRYaccParams :233-265 */
    case 8: PUT3(lerr); PUT2(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT2(prp); PUT2(pav); for (int k = 0; k < 6; k++) o[n++] = params[k]; break; /* RPYParams :268-304 */
    case 9: { static const double fake[6] = {1, 0.17, 7, 0.01, 1.2, 0.3};
              PUT3(lerr); PUT2(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT2(prp); PUT2(pav); for (int k = 0; k < 6; k++) o[n++] = fake[k]; break; } /* RPYFakeParams :307-344 */
    case 10: PUT3(lerr); PUT2(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT2(prp); PUT2(pav); break;                       /* RPY :347-382 */
    case 11: PUT3(lerr); PUT2R(rpy); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT3(acc); break;                                  /* PRYaccNoPend :385-416 (state[16:19], Q11) */
    case 13: { double r3[3] = {rpy[0], rpy[1], -hd}, Rm[9], RmT[9]; rpy2dcm(r3, Rm); m3_transpose(RmT, Rm);
               PUT3(lerr); for (int k = 0; k < 9; k++) o[n++] = RmT[k]; PUT3(lvel); PUT3(angvel); PUT2(prp); PUT2(pav); for (int k = 0; k < 6; k++) o[n++] = params[k]; break; } /* RmParams :453-489 */
    case 14: PUT3(lerr); PUT3(zvec); o[n++] = hd; PUT3(lvel); PUT3(angvel); PUT2(prp); PUT2(pav); break;                      /* Zvec :492-529 */
    default: return -2;
    }
    return n;
}

/* ------------------------------------------------------------------ Philox4x32-10 + samplers */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u01(uint32_t x) { return ((double)(x >> 8) + 0.5) * (1.0 / 16777216.0); }
static void box_muller(uint32_t x0, uint32_t x1, double *z0, double *z1) {
    double r = sqrt(-2.0 * log(u01(x0))), th = 2 * PI * u01(x1);
    *z0 = r * cos(th); *z1 = r * sin(th);
}
static double clipn(double z, double sigma) { double v = z * sigma; double lim = 2 * sigma; return v < -lim ? -lim : (v > lim ? lim : v); }
static void draw(uint32_t seed, uint32_t env, uint32_t blk, uint32_t epoch, uint32_t stream, uint32_t out[4]) {
    uint32_t ctr[4] = {blk, epoch, stream, 0}, key[2] = {seed, env};
    orc_philox4x32(ctr, key, out);
}

/* BaseDroneEnv.sample_state (:218-257); draw order :222,224,227,230,234,235,239,241 */
void orc_sample_state(const OrcResetCfg *c, uint32_t seed, uint32_t env, uint32_t rc, double *qpos, double *qvel) {
    double rpy[3] = {0, 0, c->start_pos[3]};
    for (int k = 0; k < 3; k++) qpos[k] = c->start_pos[k];
    for (int k = 0; k < 8; k++) qvel[k] = 0;
    qpos[7] = qpos[8] = 0;
    if (c->random_start_pos) {
        uint32_t x[4]; double n[4];
        draw(seed, env, 0, rc, 0, x); box_muller(x[0], x[1], &n[0], &n[1]); box_muller(x[2], x[3], &n[2], &n[3]);
        double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        draw(seed, env, 1, rc, 0, x);
        double r = c->max_pos_offset * cbrt(u01(x[0]));
        for (int k = 0; k < 3; k++) qpos[k] = c->start_pos[k] + r * (n[k] / nn);
        double yaw = PI - 2 * PI * u01(x[1]);
        box_muller(x[2], x[3], &n[0], &n[1]);
        rpy[0] = clipn(n[0], c->angle_sigma[0]); rpy[1] = clipn(n[1], c->angle_sigma[1]); rpy[2] = yaw;
        draw(seed, env, 2, rc, 0, x); box_muller(x[0], x[1], &n[0], &n[1]); box_muller(x[2], x[3], &n[2], &n[3]);
        for (int k = 0; k < 3; k++) qvel[k] = clipn(n[k], c->vel_sigma[k]);
        qvel[3] = clipn(n[3], c->ang_vel_sigma[0]);
        draw(seed, env, 3, rc, 0, x); box_muller(x[0], x[1], &n[0], &n[1]); box_muller(x[2], x[3], &n[2], &n[3]);
        qvel[4] = clipn(n[0], c->ang_vel_sigma[1]); qvel[5] = clipn(n[1], c->ang_vel_sigma[2]);
        if (c->pendulum) {
            qpos[7] = clipn(n[2], c->pend_rp_sigma[0]); qpos[8] = clipn(n[3], c->pend_rp_sigma[1]);
            draw(seed, env, 4, rc, 0, x); box_muller(x[0], x[1], &n[0], &n[1]);
            qvel[6] = clipn(n[0], c->pend_vel_sigma[0]); qvel[7] = clipn(n[1], c->pend_vel_sigma[1]);
        }
    }
    orc_rpy2quat(rpy, qpos + 3);
}
/* BaseDroneEnv.generate_drone_params (:180-216) */
void orc_sample_params(const OrcResetCfg *c, uint32_t seed, uint32_t env, uint32_t epoch, double p[6]) {
    for (int k = 0; k < 6; k++) p[k] = c->param_center[k];
    if (c->random_params) {
        uint32_t x[8];
        draw(seed, env, 0, epoch, 1, x); draw(seed, env, 1, epoch, 1, x + 4);
        for (int k = 0; k < 6; k++) {
            double w = c->param_halfwidth[k];
            p[k] = c->param_center[k] + (-w + 2 * w * u01(x[k])) * c->param_difficulty;
        }
    }
    if (!c->pendulum) { p[4] = 0; p[5] = 0; }
}

/* BaseDroneEnv.control_reference (BaseDroneEnv.py:151-172) without the joystick polling / mocap side effects.
 * axes = (x, y, z, yaw) AFTER the sign flips of :154-157, i.e. the `pert` vector of :159 (joystick.py:36 has already
 * rounded every axis to 2 decimals; that is the caller's business, not this function's).
 * ref[4] is self.reference (absolute position + yaw), start_pos[3] the clip centre (:168-169). */
void orc_control_reference(double ref[4], const double axes[4], const double start_pos[3]) {
    double pert[4], mag[4], sg[4];
    for (int k = 0; k < 4; k++) pert[k] = axes[k];
    /* np.linalg.norm of a 2-vector: sqrt(dot(x, x)) */
    const int xy_active = sqrt(pert[0] * pert[0] + pert[1] * pert[1]) > 0.2;          /* :160 */
    const int zyaw_active = sqrt(pert[2] * pert[2] + pert[3] * pert[3]) > 0.2;        /* :161 */
    for (int k = 0; k < 4; k++) {
        mag[k] = fmax(fabs(pert[k]) - 0.1, 0.0);                                      /* :162 */
        sg[k] = (pert[k] > 0) - (pert[k] < 0);                                        /* :163 */
        const int on = k < 2 ? xy_active : zyaw_active;
        ref[k] = ref[k] + 0.1 * mag[k] * sg[k] * (double)on;                          /* :164, :167 */
    }
    /* Python's float % : result takes the sign of the divisor */
    double y = fmod(ref[3] + PI, 2 * PI);
    if (y < 0) y += 2 * PI;
    ref[3] = y - PI;                                                                  /* :168 */
    const double lim[3] = {5, 5, 6};
    for (int k = 0; k < 3; k++) ref[k] = fmin(fmax(ref[k], start_pos[k] - lim[k]), start_pos[k] + lim[k]);   /* :170 */
}

/* RLlib's reset_at() round trip after vector_step (BaseDroneEnv.py:334-351), batched: every env whose
 * `truncated` flag is set gets the next draw of its own reset stream (reset_count += 1, the same bookkeeping as the
 * in-kernel reset of the CUDA path) and num_steps = 0; act / sensordata persist (Q3). */
void orc_reset_truncated(int n, int nq, int nv, const OrcResetCfg *c, uint32_t seed, uint32_t env0, uint32_t *reset_count,
                         const uint8_t *truncated, double *qpos, double *qvel, int64_t *num_steps) {
    for (int i = 0; i < n; i++) {
        if (!truncated[i]) continue;
        double qp[ORC_MAXNQ], qv[ORC_MAXNV];
        reset_count[i] += 1;
        orc_sample_state(c, seed, env0 + (uint32_t)i, reset_count[i], qp, qv);
        for (int k = 0; k < nq; k++) qpos[(size_t)nq * i + k] = qp[k];
        for (int k = 0; k < nv; k++) qvel[(size_t)nv * i + k] = qv[k];
        num_steps[i] = 0;
    }
}

/* ------------------------------------------------------------------ batched CPU vec-env step */
/* ------------------------------------------------------------------ MyBetaDist (distributions.py:6-38) on policy logits
 * alpha / beta = log(exp(clamp(x, -50, 50)) + 1) + 1 (:12-13), chunked alpha-first (:16); sample = Beta(alpha, beta) as
 * Ga / (Ga + Gb) with Marsaglia-Tsang gammas on the Philox stream (key = seed, env; counter = block, step, 2, variate);
 * deterministic = mean (:24-26); logp = sum_k log pdf(clamp(x_k, 0.01, 0.99)) (:19-22). */
static int mt_accept(double d, double c, double z, double u, double *out) {
    double t = 1.0 + c * z;
    if (!(t > 0)) return 0;
    double v = t * t * t;
    if (!(log(u) < 0.5 * z * z + d - d * v + d * log(v))) return 0;
    *out = d * v;
    return 1;
}
static double gamma_mt_retry(double d, double c, uint32_t seed, uint32_t env, uint32_t step, uint32_t vi) {
    for (uint32_t blk = 1; blk < 8; blk++) {
        uint32_t ctr[4] = {blk, step, 2u, vi}, key[2] = {seed, env}, x[4];
        orc_philox4x32(ctr, key, x);
        double z0, z1, out;
        box_muller(x[0], x[1], &z0, &z1);
        if (mt_accept(d, c, z0, u01(x[2]), &out)) return out;
        if (mt_accept(d, c, z1, u01(x[3]), &out)) return out;
    }
    return d;
}
/* one Philox block (counter block 0 of variate 2k) feeds the first attempt of both variates of action k */
static void gamma_pair(double a, double b, uint32_t seed, uint32_t env, uint32_t step, uint32_t k, double *ga, double *gb) {
    double da = a - 1.0 / 3.0, ca = 1.0 / sqrt(9.0 * da), db = b - 1.0 / 3.0, cb = 1.0 / sqrt(9.0 * db);
    uint32_t ctr[4] = {0u, step, 2u, 2u * k}, key[2] = {seed, env}, x[4];
    orc_philox4x32(ctr, key, x);
    double z0, z1;
    box_muller(x[0], x[1], &z0, &z1);
    if (!mt_accept(da, ca, z0, u01(x[2]), ga)) *ga = gamma_mt_retry(da, ca, seed, env, step, 2u * k);
    if (!mt_accept(db, cb, z1, u01(x[3]), gb)) *gb = gamma_mt_retry(db, cb, seed, env, step, 2u * k + 1u);
}
static double softplus1(double x) { x = x < -50 ? -50 : (x > 50 ? 50 : x); return log(exp(x) + 1.0) + 1.0; }
void orc_beta_policy(const double *logits, int n, int nact, uint32_t seed, uint32_t env0, uint32_t step, int deterministic,
                     double *actions, double *logp) {
    for (int i = 0; i < n; i++) {
        double lp = 0;
        for (int k = 0; k < nact; k++) {
            double a = softplus1(logits[(size_t)i * 2 * nact + k]), b = softplus1(logits[(size_t)i * 2 * nact + nact + k]), s;
            if (deterministic) s = a / (a + b);
            else {
                double ga, gb;
                gamma_pair(a, b, seed, env0 + (uint32_t)i, step, (uint32_t)k, &ga, &gb);
                s = ga / (ga + gb);
            }
            actions[(size_t)i * nact + k] = s;
            double xc = s < 1e-2 ? 1e-2 : (s > 1 - 1e-2 ? 1 - 1e-2 : s);
            lp += lgamma(a + b) - lgamma(a) - lgamma(b) + (a - 1) * log(xc) + (b - 1) * log1p(-xc);
        }
        if (logp) logp[i] = lp;
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* BaseDroneEnv.vector_step (:259-294) for n independent drones */
void orc_vector_step(int n, const OrcModel *models, int frame_skip, double *qpos, double *qvel,
                     double *act, double *sens, int64_t *num_steps, const double *actions,
                     const double *reference, int per_env_ref, int reward_id, int obs_id,
                     double max_distance, int64_t max_steps, double *obs, int obs_stride,
                     double *rewards, uint8_t *truncated, int nthreads) {
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
    for (int i = 0; i < n; i++) {
        const OrcModel *m = &models[i];
        double *qp = qpos + (size_t)m->nq * i, *qv = qvel + (size_t)m->nv * i, *a = act + 4 * (size_t)i, *sd = sens + 3 * (size_t)i;
        const double *ac = actions + 4 * (size_t)i, *ref = per_env_ref ? reference + 4 * (size_t)i : reference;
        double ctrl[4];
        for (int k = 0; k < 4; k++) ctrl[k] = 0.1 + 0.9 * ac[k];                      /* BaseDroneEnv.py:269 */
        orc_step(m, qp, qv, a, ctrl, sd, frame_skip);
        num_steps[i] += 1;
        double st[40];
        int ns = orc_drone_state(m, qp, qv, a, sd, ref, st);
        truncated[i] = (uint8_t)orc_termination(st, ref, max_distance, num_steps[i], max_steps);
        rewards[i] = orc_reward(reward_id, st, ns, ac, num_steps[i], ref, max_distance);
        if (obs) orc_obs(obs_id, st, ns, ref, obs + (size_t)obs_stride * i);
    }
}

/* `reps` consecutive vector_steps + reset_at round trips in one call (single-env configs: a Python call per step would
 * time the interpreter, not the path).  Actions cycle through `actions` [nact][n][4]. */
void orc_vector_step_repeat(int reps, int n, const OrcModel *models, int frame_skip, double *qpos, double *qvel, double *act,
                            double *sens, int64_t *num_steps, const double *actions, int nact, const double *reference,
                            int per_env_ref, int reward_id, int obs_id, double max_distance, int64_t max_steps, double *obs,
                            int obs_stride, double *rewards, uint8_t *truncated, int nthreads, const OrcResetCfg *c,
                            uint32_t seed, uint32_t env0, uint32_t *reset_count, int64_t *n_truncations) {
    for (int r = 0; r < reps; r++) {
        orc_vector_step(n, models, frame_skip, qpos, qvel, act, sens, num_steps, actions + (size_t)(r % nact) * n * 4, reference,
                        per_env_ref, reward_id, obs_id, max_distance, max_steps, obs, obs_stride, rewards, truncated, nthreads);
        for (int i = 0; i < n; i++) *n_truncations += truncated[i];
        orc_reset_truncated(n, models[0].nq, models[0].nv, c, seed, env0, reset_count, truncated, qpos, qvel, num_steps);
    }
}

/* mj_step over ONE MjData that holds n drones (drone-major qpos / qvel / act / ctrl / sensordata, BaseDroneEnv.py:367-375):
 * what `mujoco.mj_step(model, data, nstep)` does for the reference's N-drone model, single-threaded like MuJoCo. */
void orc_step_batch(int n, const OrcModel *models, double *qpos, double *qvel, double *act, const double *ctrl, double *sens, int nstep) {
    for (int i = 0; i < n; i++) {
        const OrcModel *m = &models[i];
        orc_step(m, qpos + (size_t)m->nq * i, qvel + (size_t)m->nv * i, act + 4 * (size_t)i, ctrl + 4 * (size_t)i, sens + 3 * (size_t)i, nstep);
    }
}
