/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Brute-force CPU restatement of the reference's pure-torch offroad distance:
 *   torchdrivesim/infractions.py:86-173  point_to_mesh_distance_pt
 * for points and triangles in the z = 0 plane (infractions.py:206,221 pad z with 0, so the
 * plane-projection term t is identically 0 and the "inside" distance is 0).
 * Plain fp32, no contraction (-ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>

static float edge_dist2(float px, float py, float ax, float ay, float bx, float by)
{
    float abx = bx - ax, aby = by - ay;
    float l2 = abx * abx + aby * aby;
    float t = (abx * (px - ax) + aby * (py - ay)) / (l2 + 1e-8f);
    float tt = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    float qx = ax + tt * abx, qy = ay + tt * aby;
    float dx = px - qx, dy = py - qy;
    float d = dx * dx + dy * dy;
    if (l2 <= 1e-8f) {
        float ex = px - bx, ey = py - by;
        d = ex * ex + ey * ey;
    }
    return d;
}

/* squared distance from (px,py) to one triangle, infractions.py:100-169 */
float oracle_point_tri_dist2(float px, float py, float x0, float y0, float x1, float y1, float x2, float y2)
{
    /* cross = (v2 - v0) x (v1 - v0), z component only */
    float ax = x2 - x0, ay = y2 - y0, bx = x1 - x0, by = y1 - y0;
    float cz = ax * by - ay * bx;
    float norm_normal = fabsf(cz);
    /* barycentric coordinates with p0 = v1 - v0, p1 = v2 - v0, p2 = p - v0 */
    float p0x = bx, p0y = by, p1x = ax, p1y = ay, p2x = px - x0, p2y = py - y0;
    float d00 = p0x * p0x + p0y * p0y;
    float d01 = p0x * p1x + p0y * p1y;
    float d11 = p1x * p1x + p1y * p1y;
    float d20 = p2x * p0x + p2y * p0y;
    float d21 = p2x * p1x + p2y * p1y;
    float denom = d00 * d11 - d01 * d01 + 1e-8f;
    float w1 = (d11 * d20 - d01 * d21) / denom;
    float w2 = (d00 * d21 - d01 * d20) / denom;
    float w0 = 1.0f - w1 - w2;
    int inside = (0.f <= w0) & (w0 <= 1.f) & (0.f <= w1) & (w1 <= 1.f) & (0.f <= w2) & (w2 <= 1.f);
    float area = fabsf(p0x * p1y - p0y * p1x) / 2.0f;
    int cond = inside & !(area < 5e-3f) & (norm_normal > 1e-8f);
    float e01 = edge_dist2(px, py, x0, y0, x1, y1);
    float e02 = edge_dist2(px, py, x0, y0, x2, y2);
    float e12 = edge_dist2(px, py, x1, y1, x2, y2);
    float d = fminf(fminf(e01, e02), e12);
    return cond ? 0.0f : d;
}

/* min over all faces for each point; argmin (first minimal face) optional */
void oracle_points_mesh_dist2(const float *pts /*[P][2]*/, int np, const float *verts /*[V][2]*/,
                              const int32_t *faces /*[F][3]*/, int nf, float *out_d2, int32_t *out_face)
{
    for (int p = 0; p < np; p++) {
        float px = pts[2 * p], py = pts[2 * p + 1];
        float best = INFINITY;
        int bf = -1;
        for (int f = 0; f < nf; f++) {
            const int32_t *fc = faces + 3 * f;
            float d = oracle_point_tri_dist2(px, py, verts[2 * fc[0]], verts[2 * fc[0] + 1],
                                             verts[2 * fc[1]], verts[2 * fc[1] + 1],
                                             verts[2 * fc[2]], verts[2 * fc[2] + 1]);
            if (d < best) { best = d; bf = f; }
        }
        if (best != best) best = 0.f;            /* nan_to_num, infractions.py:171 */
        if (nf == 0) best = 0.f;
        out_d2[p] = best;
        if (out_face) out_face[p] = bf;
    }
}
