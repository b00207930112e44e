// Test shim (CPU): exposes the product's host-side set-up code (qgdsolver_b200/csrc/qgd_host_setup.cpp: HostMesh::build,
// buildFaceRecords, buildLeastSquares) through a small C interface so that it can be checked against the oracle without a GPU.
// Compiled by tests/test_host_setup_cpu.py with g++ from the product sources where they lie; nothing here is shipped.
#include <cstring>

#include "qgd_internal.h"

namespace qgd { void setLastError(const std::string&) {} }

using qgd::HostMesh;

extern "C" {

void* hs_create(const qgd_mesh_desc* d)
{
    try { HostMesh* h = new HostMesh(); h->build(*d); return h; } catch (...) { return nullptr; }
}
void hs_destroy(void* p) { delete static_cast<HostMesh*>(p); }

void hs_lengths(void* p, double* hQGDf, double* hQGD)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    std::memcpy(hQGDf, h.hQGDf.data(), sizeof(double) * h.nFaces);
    std::memcpy(hQGD, h.hQGD.data(), sizeof(double) * h.nCells);
}

// vtx nFaces*4, flags nFaces, G 9*nFaces (SoA), halfDist nBnd
void hs_face_records(void* p, int reduced, int* vtx, int* flags, double* G, double* halfDist)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    std::vector<int> v, f;
    std::vector<double> g, hd;
    h.buildFaceRecords(reduced != 0, v, f, g, hd);
    std::memcpy(vtx, v.data(), sizeof(int) * v.size());
    std::memcpy(flags, f.data(), sizeof(int) * f.size());
    std::memcpy(G, g.data(), sizeof(double) * g.size());
    if (!hd.empty()) std::memcpy(halfDist, hd.data(), sizeof(double) * hd.size());
}

// point interpolation of one scalar field with the product's weights: interior points from cells, patch points from boundary values
void hs_point_interpolate(void* p, const double* cell, const double* bnd, double* out)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    for (int i = 0; i < h.nPoints; ++i) {
        double a = 0.0;
        for (int q = h.pcOff[i]; q < h.pcOff[i + 1]; ++q) a += h.pcW[q] * cell[h.pcCell[q]];
        out[i] = a;
    }
    for (size_t k = 0; k < h.patchPoints.size(); ++k) {
        double a = 0.0;
        for (int q = h.ppOff[k]; q < h.ppOff[k + 1]; ++q) a += h.ppW[q] * bnd[h.ppFace[q]];
        out[h.patchPoints[k]] = a;
    }
}

int hs_lsq_width(void* p, int opt)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    int W = 0; std::vector<int> c; std::vector<double> k; std::vector<char> d;
    h.buildLeastSquares(opt != 0, W, c, k, d);
    return W;
}
// cells W*nI, coef W*3*nI, deg nI  (column-major like the device arrays)
void hs_lsq(void* p, int opt, int* cells, double* coef, char* deg)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    int W = 0; std::vector<int> c; std::vector<double> k; std::vector<char> d;
    h.buildLeastSquares(opt != 0, W, c, k, d);
    std::memcpy(cells, c.data(), sizeof(int) * c.size());
    std::memcpy(coef, k.data(), sizeof(double) * k.size());
    std::memcpy(deg, d.data(), d.size());
}

// DIC blocks by recursive coordinate bisection (HostMesh::makePcgBlocks): block id per cell, returns the number of blocks
int hs_pcg_blocks(void* p, int target, int* out)
{
    HostMesh& h = *static_cast<HostMesh*>(p);
    h.makePcgBlocks(target);
    int nb = 0;
    for (int c = 0; c < h.nCells; ++c) { out[c] = h.pcgBlock[c]; nb = nb > out[c] + 1 ? nb : out[c] + 1; }
    return nb;
}

// cell -> faces rows (CSR) as the device ELL / tail builder consumes them
int hs_cell_faces(void* p, int* off, int* enc)
{
    const HostMesh& h = *static_cast<HostMesh*>(p);
    if (off) std::memcpy(off, h.cfOff.data(), sizeof(int) * h.cfOff.size());
    if (enc) std::memcpy(enc, h.cfEnc.data(), sizeof(int) * h.cfEnc.size());
    return (int)h.cfEnc.size();
}

}
