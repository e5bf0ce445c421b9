/* dda_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the reference viewer's DDA ray caster in DEPTH_MODE
 * (shaders/octree_dda.frag.glsl, uniforms as set by src/svviewer/octree_dda_renderer.cpp:195-211,
 * 391-397), used to check the CUDA depth-image checker (svdag-compression_b200/csrc/svb_raycast.cu)
 * pixel for pixel.  Every function cites the shader lines it follows.  float32 arithmetic throughout,
 * one rounding per operation (build with -ffp-contract=off), IEEE division and square root.
 *
 * Parity note: the reference runs this code as GLSL on an OpenGL driver, which cannot be executed
 * here; GLSL leaves normalize(), matrix products and division precision to the implementation, so
 * this restatement is "parity unpinned" against real driver output.  What it pins is our own CUDA
 * kernel: two independently written programs from the same shader must agree bit for bit.
 * Conventions chosen where GLSL is silent: mat4 * vec4 sums its four products left to right;
 * normalize(v) = v / sqrt(dot(v, v)) with dot summed left to right; 1/x is an IEEE division.
 *
 * Nothing in the product links, loads or imports this file.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y, z; } vec3;
typedef struct { int x, y, z; } ivec3;

typedef struct {
	int kind;                 /* 0 SVDAG, 1 USSVDAG, 2 SSVDAG */
	uint32_t levels;
	float bbmin[3], bbmax[3];
	float rootSide;
	const uint32_t* nodes;    /* SVDAG / USSVDAG words */
	const uint16_t* inner;    /* SSVDAG inner level data */
	const uint8_t* leaves;    /* SSVDAG 4^3 leaves, 8 bytes each */
	const uint32_t* levelOffsets;
	uint32_t innerLevels;     /* INNER_LEVELS: levels-1 (SVDAG/USSVDAG), levels-2 (SSVDAG)  (octree_dda_renderer.cpp:195-202) */
} Dag;

typedef struct {             /* traversal_status, octree_dda.frag.glsl:101-117 */
	float t_current;
	int node_index;
	uint32_t hdr;
	ivec3 mirror_mask;
	uint32_t leaf_data[2];
	ivec3 idx, local_idx;
	uint32_t child_linear_index;
	vec3 t_next_crossing, inv_ray_d;
	ivec3 delta_idx;
	int current_node_size;
	float cell_size;
	uint32_t level;
} TS;

#define MAX_STACK 40
typedef struct { int node; uint32_t hdr; int mask; } StackEnt;
typedef struct { StackEnt e[MAX_STACK]; uint32_t size; } Stack;

static int leaf_size(const Dag* d) { return d->kind == 2 ? 4 : 2; }   /* LEAF_SIZE, :156, :188, :219 */

/* :123-137 */
static void stack_push(Stack* s, int node, uint32_t hdr, ivec3 mm, uint32_t level) {
	s->e[s->size].node = node; s->e[s->size].hdr = hdr;
	s->e[s->size].mask = mm.x | (mm.y << 1) | (mm.z << 2) | (int)(level << 3);
	s->size++;
}
static void stack_pop(Stack* s, int* node, uint32_t* hdr, ivec3* mm, uint32_t* level) {
	s->size--;
	*node = s->e[s->size].node; *hdr = s->e[s->size].hdr;
	int mask = s->e[s->size].mask;
	mm->x = mask & 1; mm->y = (mask >> 1) & 1; mm->z = (mask >> 2) & 1;
	*level = (uint32_t)(mask >> 3) & 255u;
}

/* :149-152 */
static uint32_t voxel_to_linear_idx(ivec3 mm, ivec3 idx, int sz) {
	idx.x = (1 - 2 * mm.x) * idx.x + mm.x * (sz - 1);
	idx.y = (1 - 2 * mm.y) * idx.y + mm.y * (sz - 1);
	idx.z = (1 - 2 * mm.z) * idx.z + mm.z * (sz - 1);
	return (uint32_t)(idx.z + sz * (idx.y + sz * idx.x));
}

static int popc(uint32_t v) { return __builtin_popcount(v); }

/* SSVDAG helpers :284-286, :265-282 (buffer variant), :300-345 */
static int get_child_mask(uint32_t hdr, uint32_t child) { return (int)((hdr >> (child << 1)) & 3u); }

static int fetch_voxel_bit(const Dag* d, const TS* ts) {
	if (d->kind != 2) return (ts->hdr & (1u << ts->child_linear_index)) != 0;   /* :172-174, :202-204 */
	if (ts->level < d->innerLevels) return get_child_mask(ts->hdr, ts->child_linear_index) != 0;   /* :295-301 */
	uint32_t word = ts->child_linear_index & 32u;                                /* :288-293 */
	uint32_t bit = 1u << (ts->child_linear_index & 31u);
	return ((word == 0 ? ts->leaf_data[0] : ts->leaf_data[1]) & bit) != 0;
}

static void fetch_data(const Dag* d, TS* ts) {
	if (d->kind != 2) { ts->hdr = d->nodes[ts->node_index]; return; }          /* :176-178 */
	if (ts->level < d->innerLevels) ts->hdr = d->inner[(uint32_t)ts->node_index + d->levelOffsets[ts->level]];   /* :303-305 */
	else memcpy(ts->leaf_data, d->leaves + 8ull * (uint32_t)ts->node_index, 8);   /* :306-307 */
}

static void fetch_child_index_in(const Dag* d, TS* ts) {
	if (d->kind == 0) {                                                          /* :180-184 */
		int pos = popc((ts->hdr & 0xFFu) >> ts->child_linear_index);
		ts->node_index = (int)d->nodes[ts->node_index + pos];
	} else if (d->kind == 1) {                                                   /* :206-215 */
		int pos = popc((ts->hdr & 0xFFu) >> ts->child_linear_index);
		uint32_t hdr = ts->hdr;
		ts->node_index = (int)d->nodes[ts->node_index + pos];
		if (hdr & (1u << (ts->child_linear_index + 8))) ts->mirror_mask.x ^= 1;
		if (hdr & (1u << (ts->child_linear_index + 16))) ts->mirror_mask.y ^= 1;
		if (hdr & (1u << (ts->child_linear_index + 24))) ts->mirror_mask.z ^= 1;
	} else {                                                                     /* :310-345 */
		int node_offset = 1 + (int)d->levelOffsets[ts->level] + ts->node_index;
		for (uint32_t i = 7; i > ts->child_linear_index; --i) {                  /* childIndir table == this loop (renderer.cpp:533-545) */
			int cm = get_child_mask(ts->hdr, i);
			node_offset += cm < 2 ? cm : 2;
		}
		int c0 = (int)d->inner[node_offset], c1 = (int)d->inner[node_offset + 1];
		int child_mask = get_child_mask(ts->hdr, ts->child_linear_index);
		int mx = (c0 >> 13) & 1, my = (c0 >> 14) & 1, mz = (c0 >> 15) & 1;
		c0 &= -57345;
		if (child_mask > 1) {
			c0 = (int)(((uint32_t)c0 << 16) | (uint32_t)c1);
			c0 |= (child_mask & 1) << 29;
		}
		ts->mirror_mask.x ^= mx; ts->mirror_mask.y ^= my; ts->mirror_mask.z ^= mz;
		ts->node_index = c0;
	}
}

/* :350-355 */
static int in_bounds(ivec3 v, int sz) { return v.x < sz && v.y < sz && v.z < sz && 0 <= v.x && 0 <= v.y && 0 <= v.z; }

static float fmin2(float a, float b) { return a < b ? a : b; }   /* GLSL min/max: y < x ? y : x semantics on non-NaN input */
static float fmax2(float a, float b) { return a < b ? b : a; }

static ivec3 imax0(ivec3 v) { ivec3 r = { v.x > 0 ? v.x : 0, v.y > 0 ? v.y : 0, v.z > 0 ? v.z : 0 }; return r; }

/* t_next_crossing = (idx_next * cell_size - r.o) * inv_ray_d   (:386-389, :427-431, :447-451) */
static void set_crossings(TS* ts, vec3 ro) {
	ivec3 c = imax0(ts->delta_idx);
	ts->t_next_crossing.x = ((float)(ts->idx.x + c.x) * ts->cell_size - ro.x) * ts->inv_ray_d.x;
	ts->t_next_crossing.y = ((float)(ts->idx.y + c.y) * ts->cell_size - ro.y) * ts->inv_ray_d.y;
	ts->t_next_crossing.z = ((float)(ts->idx.z + c.z) * ts->cell_size - ro.z) * ts->inv_ray_d.z;
}

/* :376-390 */
static void dda_init(vec3 ro, vec3 rd, TS* ts) {
	const float voxel_eps = 1.0f / (256.f * 1024.f);
	float tt = ts->t_current + voxel_eps;
	vec3 p = { ro.x + tt * rd.x, ro.y + tt * rd.y, ro.z + tt * rd.z };
	ts->idx.x = (int)(p.x / ts->cell_size); ts->idx.y = (int)(p.y / ts->cell_size); ts->idx.z = (int)(p.z / ts->cell_size);
	set_crossings(ts, ro);
	ts->local_idx.x = ts->idx.x % 2; ts->local_idx.y = ts->idx.y % 2; ts->local_idx.z = ts->idx.z % 2;
}

/* the step mask of :398-402 / :417-421: component with the smallest t_next_crossing (x<y && x<=z, y<z && y<=x, z<x && z<=y) */
static void step_mask(const TS* ts, int m[3]) {
	vec3 t = ts->t_next_crossing;
	m[0] = (t.x < t.y) && (t.x <= t.z);
	m[1] = (t.y < t.z) && (t.y <= t.x);
	m[2] = (t.z < t.x) && (t.z <= t.y);
}

/* :397-414 */
static void dda_next(TS* ts) {
	int m[3];
	step_mask(ts, m);
	ts->idx.x += m[0] * ts->delta_idx.x; ts->idx.y += m[1] * ts->delta_idx.y; ts->idx.z += m[2] * ts->delta_idx.z;
	ts->local_idx.x += m[0] * ts->delta_idx.x; ts->local_idx.y += m[1] * ts->delta_idx.y; ts->local_idx.z += m[2] * ts->delta_idx.z;
	ts->t_current = ((float)m[0] * ts->t_next_crossing.x + (float)m[1] * ts->t_next_crossing.y) + (float)m[2] * ts->t_next_crossing.z;   /* dot() */
	ts->t_next_crossing.x += (float)m[0] * ts->cell_size * fabsf(ts->inv_ray_d.x);
	ts->t_next_crossing.y += (float)m[1] * ts->cell_size * fabsf(ts->inv_ray_d.y);
	ts->t_next_crossing.z += (float)m[2] * ts->cell_size * fabsf(ts->inv_ray_d.z);
}

/* :423-437 */
static void up_in(const Dag* d, Stack* st, vec3 ro, TS* ts) {
	uint32_t delta_level = ts->level;
	stack_pop(st, &ts->node_index, &ts->hdr, &ts->mirror_mask, &ts->level);
	delta_level -= ts->level;
	ts->idx.x >>= delta_level; ts->idx.y >>= delta_level; ts->idx.z >>= delta_level;
	ts->cell_size *= (float)(1 << delta_level);
	ts->current_node_size = ts->level < d->innerLevels ? 2 : leaf_size(d);
	ts->local_idx.x = ts->idx.x & 1; ts->local_idx.y = ts->idx.y & 1; ts->local_idx.z = ts->idx.z & 1;
	set_crossings(ts, ro);
}

/* :439-455 */
static void go_down_one_level(vec3 ro, vec3 rd, TS* ts) {
	ts->level++;
	ts->cell_size *= 0.5f;
	vec3 p = { ro.x + ts->t_current * rd.x, ro.y + ts->t_current * rd.y, ro.z + ts->t_current * rd.z };
	vec3 pc = { (float)(ts->idx.x * 2 + 1) * ts->cell_size, (float)(ts->idx.y * 2 + 1) * ts->cell_size, (float)(ts->idx.z * 2 + 1) * ts->cell_size };
	ts->idx.x = ts->idx.x * 2 + (pc.x < p.x); ts->idx.y = ts->idx.y * 2 + (pc.y < p.y); ts->idx.z = ts->idx.z * 2 + (pc.z < p.z);
	set_crossings(ts, ro);
	int msk = ts->current_node_size - 1;
	ts->local_idx.x = ts->idx.x & msk; ts->local_idx.y = ts->idx.y & msk; ts->local_idx.z = ts->idx.z & msk;
}

/* :457-481 */
static void down_in(const Dag* d, Stack* st, vec3 ro, vec3 rd, TS* ts) {
	int m[3];
	step_mask(ts, m);
	ivec3 nx = { ts->local_idx.x + m[0] * ts->delta_idx.x, ts->local_idx.y + m[1] * ts->delta_idx.y, ts->local_idx.z + m[2] * ts->delta_idx.z };
	if (in_bounds(nx, 2)) stack_push(st, ts->node_index, ts->hdr, ts->mirror_mask, ts->level);
	fetch_child_index_in(d, ts);
	go_down_one_level(ro, rd, ts);
	if (ts->level == d->innerLevels) {
		ts->current_node_size = leaf_size(d);
		int voxel_count = leaf_size(d) / 2;
		while (voxel_count > 1) { go_down_one_level(ro, rd, ts); voxel_count >>= 1; }
	}
}

static float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

/* trace_ray :539-585 with transform_ray :484-512, init :515-531.  out = (t, level, iterations) or t < 0 codes */
static void trace_ray(const Dag* d, vec3 ro, vec3 rd, float tmin, float tmax, float projection_factor, uint32_t maxIters, uint32_t drawLevel, float out[3]) {
	const float rootHalfSide = d->rootSide / 2.0f;                 /* getHalfSide(0), encoded_octree.hpp:68-70 */
	const float epsilon = 1e-4f;
	vec3 sc = { (d->bbmin[0] + d->bbmax[0]) * 0.5f, (d->bbmin[1] + d->bbmax[1]) * 0.5f, (d->bbmin[2] + d->bbmax[2]) * 0.5f };   /* box.center() */
	vec3 sr = { sgn(rd.x), sgn(rd.y), sgn(rd.z) };
	const float scale = 1.0f / (2.0f * rootHalfSide);
	vec3 omin = { sc.x - rootHalfSide, sc.y - rootHalfSide, sc.z - rootHalfSide };
	ro.x = (ro.x - omin.x) * scale; ro.y = (ro.y - omin.y) * scale; ro.z = (ro.z - omin.z) * scale;
	tmin *= scale; tmax *= scale;
	if (rd.x * sr.x < epsilon) rd.x = sr.x * epsilon;
	if (rd.y * sr.y < epsilon) rd.y = sr.y * epsilon;
	if (rd.z * sr.z < epsilon) rd.z = sr.z * epsilon;
	vec3 cmin = { (d->bbmin[0] - omin.x) * scale, (d->bbmin[1] - omin.y) * scale, (d->bbmin[2] - omin.z) * scale };
	vec3 cmax = { (d->bbmax[0] - omin.x) * scale, (d->bbmax[1] - omin.y) * scale, (d->bbmax[2] - omin.z) * scale };
	/* intersectAABB :357-366 */
	float t1x = (cmin.x - ro.x) / rd.x, t1y = (cmin.y - ro.y) / rd.y, t1z = (cmin.z - ro.z) / rd.z;
	float t2x = (cmax.x - ro.x) / rd.x, t2y = (cmax.y - ro.y) / rd.y, t2z = (cmax.z - ro.z) / rd.z;
	float lox = fmin2(t1x, t2x), loy = fmin2(t1y, t2y), loz = fmin2(t1z, t2z);
	float hix = fmax2(t1x, t2x), hiy = fmax2(t1y, t2y), hiz = fmax2(t1z, t2z);
	float tix = fmax2(fmax2(lox, 0.0f), fmax2(loy, loz)), tiy = fmin2(hix, fmin2(hiy, hiz));
	tmin = fmax2(tix, tmin + 1e-10f);
	tmax = fmin2(tiy, tmax);
	if (!(tix < tiy)) { out[0] = -4.f; out[1] = 0; out[2] = 0; return; }

	const float oscale = 2.0f * rootHalfSide;
	TS ts;
	Stack st; st.size = 0;
	memset(&ts, 0, sizeof(ts));
	ts.t_current = tmin;
	ts.inv_ray_d.x = 1.0f / rd.x; ts.inv_ray_d.y = 1.0f / rd.y; ts.inv_ray_d.z = 1.0f / rd.z;
	ts.delta_idx.x = (int)sgn(rd.x); ts.delta_idx.y = (int)sgn(rd.y); ts.delta_idx.z = (int)sgn(rd.z);
	ts.level = 0; ts.cell_size = 0.5f;
	dda_init(ro, rd, &ts);
	ts.current_node_size = 2;
	ts.node_index = 0;
	fetch_data(d, &ts);
	ts.child_linear_index = voxel_to_linear_idx(ts.mirror_mask, ts.local_idx, ts.current_node_size);

	uint32_t it = 0;
	const uint32_t max_level = d->innerLevels < drawLevel - 1 ? d->innerLevels : drawLevel - 1;
	do {
		int full = fetch_voxel_bit(d, &ts);
		if (!full) {
			dda_next(&ts);
			if (!in_bounds(ts.local_idx, ts.current_node_size)) {
				if (st.size == 0) { out[0] = -1.f; out[1] = 0; out[2] = (float)it; return; }
				up_in(d, &st, ro, &ts);
			}
		} else {
			int hit = ts.level >= max_level || (ts.cell_size * projection_factor) < ts.t_current;   /* resolution_ok :368-370 */
			if (hit) { out[0] = ts.t_current * oscale; out[1] = (float)ts.level; out[2] = (float)it; return; }
			down_in(d, &st, ro, rd, &ts);
			fetch_data(d, &ts);
		}
		ts.child_linear_index = voxel_to_linear_idx(ts.mirror_mask, ts.local_idx, ts.current_node_size);
		++it;
	} while (ts.t_current < tmax && it < maxIters);
	out[0] = it >= maxIters ? -3.f : -2.f; out[1] = 0; out[2] = (float)it;
}

static vec3 normalize3(vec3 v) {
	float l = sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
	vec3 r = { v.x / l, v.y / l, v.z / l };
	return r;
}
/* column-major mat4 * vec4 (GLSL), products summed left to right */
static void mat4_mul(const float* m, const float v[4], float o[4]) {
	for (int i = 0; i < 4; ++i) o[i] = ((m[0 + i] * v[0] + m[4 + i] * v[1]) + m[8 + i] * v[2]) + m[12 + i] * v[3];
}

/* computeCameraRay :592-608 */
static void camera_ray(const float* viewInv, const float* projInv, float sx, float sy, vec3* ro, vec3* rd) {
	float s0[4] = { sx, sy, 0.f, 1.f }, s1[4] = { sx, sy, 1.f, 1.f }, w0[4], w1[4];
	mat4_mul(projInv, s0, w0); mat4_mul(projInv, s1, w1);
	vec3 a = { w0[0] / w0[3], w0[1] / w0[3], w0[2] / w0[3] }, b = { w1[0] / w1[3], w1[1] / w1[3], w1[2] / w1[3] };
	vec3 dd = { b.x - a.x, b.y - a.y, b.z - a.z };
	dd = normalize3(dd);
	float o4[4] = { 0.f, 0.f, 0.f, 1.f }, e4[4] = { dd.x, dd.y, dd.z, 1.f }, op[4], ep[4];
	mat4_mul(viewInv, o4, op); mat4_mul(viewInv, e4, ep);
	ro->x = op[0]; ro->y = op[1]; ro->z = op[2];
	vec3 d2 = { ep[0] - op[0], ep[1] - op[1], ep[2] - op[2] };
	*rd = normalize3(d2);
}

/* file: the bytes svbuilder wrote (encoded_svdag.cpp:76-103, encoded_ussvdag.cpp:60-84, encoded_ssvdag.cpp:84-117) */
static int parse(const uint8_t* f, uint64_t size, int kind, Dag* d) {
	if (size < 36) return -1;
	memset(d, 0, sizeof(*d));
	d->kind = kind;
	memcpy(d->bbmin, f, 12); memcpy(d->bbmax, f + 12, 12);
	memcpy(&d->rootSide, f + 24, 4); memcpy(&d->levels, f + 28, 4);
	if (kind != 2) {
		uint32_t count; memcpy(&count, f + 40, 4);
		if (size < 44 + 4ull * count) return -1;
		d->nodes = (const uint32_t*)(f + 44);
		d->innerLevels = d->levels - 1;
	} else {
		uint64_t o = 36;
		uint32_t n; memcpy(&n, f + o, 4); o += 4; d->inner = (const uint16_t*)(f + o); o += 2ull * n;
		memcpy(&n, f + o, 4); o += 4; d->leaves = f + o; o += n;
		memcpy(&n, f + o, 4); o += 4; d->levelOffsets = (const uint32_t*)(f + o);
		d->innerLevels = d->levels - 2;
	}
	return 0;
}

/* DEPTH_MODE main() :846-858: out[(y*w + x)*3 ..] = (t, level, iterations) when t > 0, else left at 0 ("discard") */
int dda_oracle_render(const uint8_t* file, uint64_t size, int kind, const float* viewInv, const float* projInv, uint32_t w, uint32_t h,
                      uint32_t maxIters, uint32_t drawLevel, float projectionFactor, float* out) {
	Dag d;
	if (parse(file, size, kind, &d)) return -1;
	if (drawLevel == 0) drawLevel = d.levels;                      /* _drawLevel = getNLevels(), octree_dda_renderer.cpp:222 */
	memset(out, 0, sizeof(float) * 3ull * w * h);
	for (uint32_t y = 0; y < h; ++y)
		for (uint32_t x = 0; x < w; ++x) {
			float sx = (((float)x + 0.5f) / (float)w) * 2.0f - 1.0f, sy = (((float)y + 0.5f) / (float)h) * 2.0f - 1.0f;
			vec3 ro, rd;
			camera_ray(viewInv, projInv, sx, sy, &ro, &rd);
			float r[3];
			trace_ray(&d, ro, rd, 0.f, 1e30f, projectionFactor, maxIters, drawLevel, r);
			if (r[0] > 0.f) { float* o = out + 3ull * ((uint64_t)y * w + x); o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; }
		}
	return 0;
}
