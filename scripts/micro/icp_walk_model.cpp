// Host model of the ICP grid walk (csrc/opb_icp.cu: grid_nearest, unseeded) that counts what one query costs: rows whose
// extents are read, points tested, rings reached, dependent round trips (one per row for the extents, one more when the row
// holds points).  Developer tool for the load-balance question of DESIGN.md §4; not part of the product or of the tests.
//   g++ -O2 -o /tmp/icp_walk_model scripts/micro/icp_walk_model.cpp
//   /tmp/icp_walk_model src.f32 tgt.f32 n_src n_tgt radius guard_cells out.i32     (out: n_src x 4 ints: rows, points, rings, hops)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Grid { float origin[3], h, inv_h; int dim[3]; };
static std::vector<unsigned> cell_start;
static std::vector<float> sorted; // x y z per point, cell order
struct Count { int rows = 0, points = 0, rings = 0, hops = 0; };

static void scan(const Grid &g, int xa, int xb, int cy, int cz, float qx, float qy, float qz, float &d1, Count &c)
{
    xa = std::max(xa, 0); xb = std::min(xb, g.dim[0] - 1);
    if (xa > xb || cy < 0 || cy >= g.dim[1] || cz < 0 || cz >= g.dim[2]) return;
    const size_t row = (size_t)g.dim[0] * ((size_t)cy + (size_t)g.dim[1] * (size_t)cz);
    const unsigned s = cell_start[row + xa], e = cell_start[row + xb + 1];
    ++c.rows; ++c.hops;
    if (e > s) ++c.hops;
    for (unsigned k = s; k < e; ++k)
    {
        const float dx = qx - sorted[3 * k], dy = qy - sorted[3 * k + 1], dz = qz - sorted[3 * k + 2];
        d1 = std::min(d1, dx * dx + dy * dy + dz * dz);
        ++c.points;
    }
}
struct Pruner
{
    float fx, ay, az, inv_h2, slack;
    float gap(float a, int d) const { return d == 0 ? 0.0f : std::max((d < 0 ? a : 1.0f - a) + (float)(std::abs(d) - 1) - slack, 0.0f); }
    bool interval(int dy, int dz, float eb, int &xlo, int &xhi) const
    {
        const float gy = gap(ay, dy), gz = gap(az, dz), rem = eb * inv_h2 - (gy * gy + gz * gz);
        if (!(rem >= 0.0f)) return false;
        const float half = std::sqrt(rem) + slack;
        xlo = (int)std::floor(fx - half); xhi = (int)std::floor(fx + half);
        return true;
    }
};
static float bound(float bd, float guard, float cap2) { const float r = std::sqrt(bd) + guard; return std::min(r * r, cap2); }

int main(int argc, char **argv)
{
    if (argc < 8) return 1;
    const int ns = atoi(argv[3]), nt = atoi(argv[4]);
    const float radius = (float)atof(argv[5]), guard_cells = (float)atof(argv[6]);
    std::vector<float> src(3 * (size_t)ns), tgt(3 * (size_t)nt);
    FILE *f = fopen(argv[1], "rb"); if (!f || fread(src.data(), 4, src.size(), f) != src.size()) return 2; fclose(f);
    f = fopen(argv[2], "rb"); if (!f || fread(tgt.data(), 4, tgt.size(), f) != tgt.size()) return 2; fclose(f);
    Grid g;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f}, ext[3];
    for (int i = 0; i < nt; ++i) for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], tgt[3 * i + a]); hi[a] = std::max(hi[a], tgt[3 * i + a]); }
    for (int a = 0; a < 3; ++a) ext[a] = hi[a] - lo[a];
    const float emax = std::max(ext[0], std::max(ext[1], ext[2])), emin = std::min(ext[0], std::min(ext[1], ext[2]));
    const float emid = ext[0] + ext[1] + ext[2] - emax - emin;
    float h = 2.0f * std::sqrt(std::max(emax * emid, 1e-12f) / (float)nt);
    const double max_cells = std::max(16.0 * nt, (double)(1u << 20));
    for (;;) { double cells = 1; for (int a = 0; a < 3; ++a) cells *= std::floor((double)ext[a] / h) + 1.0; if (cells <= max_cells) break; h *= 1.26f; }
    g.h = h; g.inv_h = 1.0f / h;
    for (int a = 0; a < 3; ++a) { g.origin[a] = lo[a]; g.dim[a] = (int)std::floor(ext[a] / h) + 1; }
    const size_t n_cells = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
    std::vector<unsigned> cell(nt);
    cell_start.assign(n_cells + 1, 0);
    for (int i = 0; i < nt; ++i)
    {
        size_t c[3];
        for (int a = 0; a < 3; ++a) c[a] = (size_t)std::min(std::max((int)std::floor((tgt[3 * i + a] - g.origin[a]) * g.inv_h), 0), g.dim[a] - 1);
        cell[i] = (unsigned)(c[0] + g.dim[0] * (c[1] + g.dim[1] * c[2]));
        ++cell_start[cell[i] + 1];
    }
    for (size_t c = 0; c < n_cells; ++c) cell_start[c + 1] += cell_start[c];
    sorted.resize(3 * (size_t)nt);
    { std::vector<unsigned> cur(cell_start.begin(), cell_start.end() - 1);
      for (int i = 0; i < nt; ++i) { const unsigned k = cur[cell[i]]++; for (int a = 0; a < 3; ++a) sorted[3 * k + a] = tgt[3 * i + a]; } }
    fprintf(stderr, "grid h %.5f dims %d %d %d cells %zu\n", h, g.dim[0], g.dim[1], g.dim[2], n_cells);
    const float guard = guard_cells * h, cap_g = radius + guard, cap2 = cap_g * cap_g, inf = INFINITY;
    std::vector<int> out(4 * (size_t)ns);
    for (int i = 0; i < ns; ++i)
    {
        const float qx = src[3 * i], qy = src[3 * i + 1], qz = src[3 * i + 2];
        const float fx = (qx - g.origin[0]) * g.inv_h, fy = (qy - g.origin[1]) * g.inv_h, fz = (qz - g.origin[2]) * g.inv_h;
        const int hx = (int)std::floor(fx), hy = (int)std::floor(fy), hz = (int)std::floor(fz);
        Pruner pr{fx, fy - hy, fz - hz, g.inv_h * g.inv_h, 1e-3f};
        Count c; float d1 = inf; int xlo, xhi;
        scan(g, hx, hx, hy, hz, qx, qy, qz, d1, c);
        if (pr.interval(0, 0, bound(d1, guard, cap2), xlo, xhi))
        {
            if (xlo < hx) scan(g, xlo, hx - 1, hy, hz, qx, qy, qz, d1, c);
            if (xhi > hx) scan(g, hx + 1, xhi, hy, hz, qx, qy, qz, d1, c);
        }
        const float m_yz = std::min(std::min(pr.ay, 1.0f - pr.ay), std::min(pr.az, 1.0f - pr.az));
        const float radius_cells = cap_g * g.inv_h;
        const int r_max = (int)std::ceil(radius_cells) + 1;
        for (int r = 1; r <= r_max; ++r)
        {
            const float covered = (float)(r - 1) + m_yz - pr.slack;
            float eb = bound(d1, guard, cap2);
            if (covered > 0.0f && d1 < inf && eb * pr.inv_h2 <= covered * covered) break;
            if (covered > radius_cells) break;
            c.rings = r;
            if (r == 1)
            {
                static const int DY[8] = {-1, 1, 0, 0, -1, -1, 1, 1}, DZ[8] = {0, 0, -1, 1, -1, 1, -1, 1};
                unsigned mask = 0;
                for (int b = 0; b < 8; ++b) if (pr.interval(DY[b], DZ[b], eb, xlo, xhi)) mask |= 1u << b;
                for (int b = 0; b < 8; ++b)
                    if ((mask >> b & 1) && pr.interval(DY[b], DZ[b], eb, xlo, xhi)) { scan(g, xlo, xhi, hy + DY[b], hz + DZ[b], qx, qy, qz, d1, c); eb = bound(d1, guard, cap2); }
            }
            else
                for (int dz = -r; dz <= r; ++dz)
                    for (int dy = -r; dy <= r; dy += (std::abs(dz) == r ? 1 : 2 * r))
                        if (pr.interval(dy, dz, eb, xlo, xhi)) { scan(g, xlo, xhi, hy + dy, hz + dz, qx, qy, qz, d1, c); eb = bound(d1, guard, cap2); }
        }
        out[4 * i] = c.rows; out[4 * i + 1] = c.points; out[4 * i + 2] = c.rings; out[4 * i + 3] = c.hops;
    }
    f = fopen(argv[7], "wb"); fwrite(out.data(), 4, out.size(), f); fclose(f);
    return 0;
}
