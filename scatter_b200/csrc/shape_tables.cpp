// Element-independent integration tables: shape functions N and natural derivatives dN at the Gauss points,
// evaluated once on the host and staged into shared memory by the assembly kernel.
//
// Follows the reference's definitions (gmsh node ordering):
//   shape functions   scatter/element_types.py:52-104 (hexa8) :155-254 (hexa20) :303-333 (quad4) :382-424 (quad8)
//                     :475-501 (tri3) :552-587 (tri6) :646-683 (tetra4) :742-802 (tetra10)
//   Gauss tables      scatter/discretisation.py:436-497; point order u outer / w inner (:32-38, :303-306)
// quad8 keeps the reference's plain bilinear corner functions and 1/2-scaled mid-side functions (parity).
#include <cmath>
#include "common.h"

namespace {

const int kNne[8] = {3, 6, 4, 8, 4, 10, 8, 20};
const int kDim[8] = {2, 2, 2, 2, 3, 3, 3, 3};

const double HEXC[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
const int HEX20_EDGE[12][2] = {{0, 1}, {0, 3}, {0, 4}, {1, 2}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 7}, {5, 6}, {6, 7}};
const double QUADC[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};

void eval(int et, const double* xi, double* N, double* dN) {
    switch (et) {
        case SC_HEXA8: {
            for (int a = 0; a < 8; ++a) {
                double f[3];
                for (int d = 0; d < 3; ++d) f[d] = 1.0 + HEXC[a][d] * xi[d];
                N[a] = f[0] * f[1] * f[2] / 8.0;
                dN[a * 3 + 0] = HEXC[a][0] * f[1] * f[2] / 8.0;
                dN[a * 3 + 1] = f[0] * HEXC[a][1] * f[2] / 8.0;
                dN[a * 3 + 2] = f[0] * f[1] * HEXC[a][2] / 8.0;
            }
            break;
        }
        case SC_HEXA20: {
            for (int a = 0; a < 8; ++a) {
                double f[3], s = -2.0;
                for (int d = 0; d < 3; ++d) { f[d] = 1.0 + HEXC[a][d] * xi[d]; s += HEXC[a][d] * xi[d]; }
                N[a] = f[0] * f[1] * f[2] * s / 8.0;
                for (int d = 0; d < 3; ++d) {
                    int o1 = (d + 1) % 3, o2 = (d + 2) % 3;
                    dN[a * 3 + d] = HEXC[a][d] * f[o1] * f[o2] * (s + f[d]) / 8.0;
                }
            }
            for (int k = 0; k < 12; ++k) {
                const double* ca = HEXC[HEX20_EDGE[k][0]];
                const double* cb = HEXC[HEX20_EDGE[k][1]];
                double mid[3];
                int e = 0;
                for (int d = 0; d < 3; ++d) { mid[d] = 0.5 * (ca[d] + cb[d]); if (mid[d] == 0.0) e = d; }
                int o1 = (e + 1) % 3, o2 = (e + 2) % 3;
                double g1 = 1.0 + mid[o1] * xi[o1], g2 = 1.0 + mid[o2] * xi[o2], q = 1.0 - xi[e] * xi[e];
                int a = 8 + k;
                N[a] = q * g1 * g2 / 4.0;
                dN[a * 3 + e] = -2.0 * xi[e] * g1 * g2 / 4.0;
                dN[a * 3 + o1] = q * mid[o1] * g2 / 4.0;
                dN[a * 3 + o2] = q * g1 * mid[o2] / 4.0;
            }
            break;
        }
        case SC_QUAD4:
        case SC_QUAD8: {
            double u = xi[0], v = xi[1];
            for (int a = 0; a < 4; ++a) {
                double f0 = 1.0 + QUADC[a][0] * u, f1 = 1.0 + QUADC[a][1] * v;
                N[a] = f0 * f1 / 4.0;
                dN[a * 2 + 0] = QUADC[a][0] * f1 / 4.0;
                dN[a * 2 + 1] = f0 * QUADC[a][1] / 4.0;
            }
            if (et == SC_QUAD8) {
                N[4] = 0.5 * (1 - u * u) * (1 - v); dN[8] = -u * (1 - v);          dN[9] = -0.5 * (1 - u * u);
                N[5] = 0.5 * (1 + u) * (1 - v * v); dN[10] = 0.5 * (1 - v * v);    dN[11] = -v * (1 + u);
                N[6] = 0.5 * (1 - u * u) * (1 + v); dN[12] = -u * (1 + v);         dN[13] = 0.5 * (1 - u * u);
                N[7] = 0.5 * (1 - u) * (1 - v * v); dN[14] = -0.5 * (1 - v * v);   dN[15] = -v * (1 - u);
            }
            break;
        }
        case SC_TRI3: {
            double u = xi[0], v = xi[1];
            N[0] = 1 - u - v; N[1] = u; N[2] = v;
            const double d[6] = {-1, -1, 1, 0, 0, 1};
            for (int i = 0; i < 6; ++i) dN[i] = d[i];
            break;
        }
        case SC_TRI6: {
            double u = xi[0], v = xi[1], L = 1 - u - v;
            N[0] = (2 * L - 1) * L; N[1] = (2 * u - 1) * u; N[2] = (2 * v - 1) * v;
            N[3] = 4 * L * u; N[4] = 4 * u * v; N[5] = 4 * L * v;
            const double d[12] = {1 - 4 * L, 1 - 4 * L, 4 * u - 1, 0, 0, 4 * v - 1,
                                  4 * L - 4 * u, -4 * u, 4 * v, 4 * u, -4 * v, 4 * L - 4 * v};
            for (int i = 0; i < 12; ++i) dN[i] = d[i];
            break;
        }
        case SC_TETRA4: {
            double u = xi[0], v = xi[1], w = xi[2];
            N[0] = 1 - u - v - w; N[1] = u; N[2] = v; N[3] = w;
            const double d[12] = {-1, -1, -1, 1, 0, 0, 0, 1, 0, 0, 0, 1};
            for (int i = 0; i < 12; ++i) dN[i] = d[i];
            break;
        }
        case SC_TETRA10: {
            double u = xi[0], v = xi[1], w = xi[2], x = 1 - u - v - w, d0 = 1 - 4 * x;
            N[0] = (2 * x - 1) * x; N[1] = (2 * u - 1) * u; N[2] = (2 * v - 1) * v; N[3] = (2 * w - 1) * w;
            N[4] = 4 * u * x; N[5] = 4 * u * v; N[6] = 4 * v * x; N[7] = 4 * w * x; N[8] = 4 * v * w; N[9] = 4 * u * w;
            const double d[30] = {d0, d0, d0, 4 * u - 1, 0, 0, 0, 4 * v - 1, 0, 0, 0, 4 * w - 1,
                                  4 * x - 4 * u, -4 * u, -4 * u, 4 * v, 4 * u, 0, -4 * v, 4 * x - 4 * v, -4 * v,
                                  -4 * w, -4 * w, 4 * x - 4 * w, 0, 4 * w, 4 * v, 4 * w, 0, 4 * u};
            for (int i = 0; i < 30; ++i) dN[i] = d[i];
            break;
        }
    }
}

bool line_rule(int n, double* x, double* w) {
    if (n == 1) { x[0] = 0.0; w[0] = 2.0; return true; }
    if (n == 2) { double a = std::sqrt(1.0 / 3.0); x[0] = -a; x[1] = a; w[0] = w[1] = 1.0; return true; }
    if (n == 3) {
        double a = std::sqrt(3.0 / 5.0);
        x[0] = -a; x[1] = 0.0; x[2] = a; w[0] = 5.0 / 9.0; w[1] = 8.0 / 9.0; w[2] = 5.0 / 9.0;
        return true;
    }
    return false;
}

}  // namespace

int sc_elem_nne(int et) { return (et >= 0 && et < 8) ? kNne[et] : 0; }
int sc_elem_dim(int et) { return (et >= 0 && et < 8) ? kDim[et] : 0; }

bool sc_make_shape_table(int et, int order, ShapeTable& t, std::string& err) {
    if (et < 0 || et > 7) { err = "element type not supported"; return false; }
    t.nne = kNne[et];
    t.dim = kDim[et];
    std::vector<double> pts;   // [ngp][dim]
    t.w.clear();
    const bool tensor = (et == SC_QUAD4 || et == SC_QUAD8 || et == SC_HEXA8 || et == SC_HEXA20);
    if (tensor) {
        double x[3], w[3];
        if (!line_rule(order, x, w)) { err = "ERROR: integration order not supported"; return false; }
        if (t.dim == 3) {
            for (int i = 0; i < order; ++i)
                for (int j = 0; j < order; ++j)
                    for (int k = 0; k < order; ++k) {
                        pts.push_back(x[i]); pts.push_back(x[j]); pts.push_back(x[k]);
                        t.w.push_back(w[i] * w[j] * w[k]);
                    }
        } else {
            for (int i = 0; i < order; ++i)
                for (int j = 0; j < order; ++j) {
                    pts.push_back(x[i]); pts.push_back(x[j]);
                    t.w.push_back(w[i] * w[j]);
                }
        }
    } else if (t.dim == 2) {   // triangles
        if (order == 1) { pts = {1.0 / 3, 1.0 / 3}; t.w = {0.5}; }
        else if (order == 2) { pts = {1.0 / 6, 1.0 / 6, 2.0 / 3, 1.0 / 6, 1.0 / 6, 2.0 / 3}; t.w = {1.0 / 6, 1.0 / 6, 1.0 / 6}; }
        else if (order == 3) {
            pts = {1.0 / 3, 1.0 / 3, 1.0 / 5, 1.0 / 5, 3.0 / 5, 1.0 / 5, 1.0 / 5, 3.0 / 5};
            t.w = {-27.0 / 96, 25.0 / 96, 25.0 / 96, 25.0 / 96};
        } else { err = "ERROR: integration order not supported"; return false; }
    } else {                   // tetrahedra
        if (order == 1) { pts = {0.25, 0.25, 0.25}; t.w = {1.0 / 6}; }
        else if (order == 2) {
            double a = 0.25 - std::sqrt(5.0) / 20.0, b = 0.25 + 3.0 * std::sqrt(5.0) / 20.0;
            pts = {a, a, a, a, a, b, a, b, a, b, a, a};
            t.w = {1.0 / 24, 1.0 / 24, 1.0 / 24, 1.0 / 24};
        } else { err = "ERROR: integration order not supported for type tetra"; return false; }
    }
    t.ngp = (int)t.w.size();
    t.N.assign((size_t)t.ngp * t.nne, 0.0);
    t.dN.assign((size_t)t.ngp * t.nne * t.dim, 0.0);
    for (int g = 0; g < t.ngp; ++g) eval(et, &pts[(size_t)g * t.dim], &t.N[(size_t)g * t.nne], &t.dN[(size_t)g * t.nne * t.dim]);
    return true;
}
