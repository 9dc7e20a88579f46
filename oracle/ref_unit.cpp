// oracle/ref_unit.cpp — extern "C" doorways onto individual functions of the UNMODIFIED reference
// (linked against oracle/_ref/libmcut_ref.so) so python can feed them vectors and record what the
// reference answers.  Used only by tests/golden/make_golden.py and the ref-gated tests.
// TEST INFRASTRUCTURE ONLY.
#include <vector>

#include "mcut/mcut.h"
#include "mcut/internal/bvh.h"
#include "mcut/internal/math.h"
#include "mcut/internal/hmesh.h"
#include "mcut/internal/tpool.h"

extern unsigned int morton3D(float x, float y, float z); // source/bvh.cpp:206
extern int get_ostensibly_implicit_bvh_size(const int t); // source/bvh.cpp:131
extern bool calculate_vertex_parameters(double& quantization_multiplier, vec3_<double>& pre_quantization_translation,
    vec3_<double>& srcmesh_cutmesh_com, vec3_<double>& srcmesh_bboxmin, vec3_<double>& srcmesh_bboxmax,
    vec3_<double>& cutmesh_bboxmin, vec3_<double>& cutmesh_bboxmax, McFlags dispatchFlags, const void* pSrcMeshVertices,
    McUint32 numSrcMeshVertices, const void* pCutMeshVertices, McUint32 numCutMeshVertices); // source/preproc.cpp:2124

static std::vector<vec3> to_vec(const double* v, int n)
{
    std::vector<vec3> out;
    for (int i = 0; i < n; ++i) out.push_back(vec3(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
    return out;
}

extern int find_connected_components(thread_pool& scheduler, std::vector<int>& fccmap, const hmesh_t& mesh,
    std::vector<int>& cc_to_vertex_count, std::vector<int>& cc_to_face_count); // source/kernel.cpp:235
extern bool mesh_is_closed(const hmesh_t& mesh); // source/preproc.cpp:1957

extern double calculate_signed_solid_angle(const vec3_<double>& a, const vec3_<double>& b, const vec3_<double>& c,
    const vec3_<double>& query, const double multiplier); // source/preproc.cpp:1650
extern double calculate_signed_solid_angle(const vec3_<double>& a, const vec3_<double>& b, const vec3_<double>& c,
    const vec3_<double>& d, const vec3_<double>& query, const double multiplier); // source/preproc.cpp:1702

extern "C" {

double ref_solid_angle_tri(const double* a, const double* b, const double* c, const double* q)
{
    return calculate_signed_solid_angle(vec3_<double>(a[0], a[1], a[2]), vec3_<double>(b[0], b[1], b[2]), vec3_<double>(c[0], c[1], c[2]),
        vec3_<double>(q[0], q[1], q[2]), 1.0);
}
double ref_solid_angle_quad(const double* a, const double* b, const double* c, const double* d, const double* q)
{
    return calculate_signed_solid_angle(vec3_<double>(a[0], a[1], a[2]), vec3_<double>(b[0], b[1], b[2]), vec3_<double>(c[0], c[1], c[2]),
        vec3_<double>(d[0], d[1], d[2]), vec3_<double>(q[0], q[1], q[2]), 1.0);
}

// find_connected_components + mesh_is_closed on a mesh given as arrays (the half-edge mesh is built the reference's way:
// add_vertex / add_face in order).  Returns the number of components, -1 if a face could not be added.
int ref_validate(int nv, const unsigned* face_off, const unsigned* face_vtx, int nf, int* fccmap, int* cc_vertex_count,
    int* cc_face_count, int* is_closed)
{
    hmesh_t m;
    for (int v = 0; v < nv; ++v) m.add_vertex(vec3((double)v, 0.0, 0.0));
    std::vector<vd_t> fv;
    for (int f = 0; f < nf; ++f) {
        fv.clear();
        for (unsigned h = face_off[f]; h < face_off[f + 1]; ++h) fv.push_back(vd_t(face_vtx[h]));
        if (m.add_face(fv) == hmesh_t::null_face()) return -1;
    }
    // zero helper threads: with helpers the reference's per-component face count is a data race (`cc_to_face_count[id] += 1`
    // from several pool threads without atomics, kernel.cpp:330-352) and comes out short on meshes above 2048 faces
    thread_pool pool(0, 1);
    std::vector<int> map, cv, cf;
    const int n = find_connected_components(pool, map, m, cv, cf);
    for (int f = 0; f < nf; ++f) fccmap[f] = map[f];
    for (int c = 0; c < n; ++c) {
        cc_vertex_count[c] = cv[c];
        cc_face_count[c] = cf[c];
    }
    *is_closed = mesh_is_closed(m) ? 1 : 0;
    return n;
}

unsigned ref_morton3D(float x, float y, float z) { return morton3D(x, y, z); }
int ref_oibvh_size(int t) { return get_ostensibly_implicit_bvh_size(t); }

double ref_orient3d(const double* a, const double* b, const double* c, const double* d) { return ::orient3d(a, b, c, d); }
double ref_orient2d(const double* a, const double* b, const double* c) { return ::orient2d(a, b, c); }

int ref_plane_coefficients(const double* verts, int n, double* normal, double* d)
{
    std::vector<vec3> v = to_vec(verts, n);
    vec3 nrm; // value-initialised like the reference's unordered_map::operator[] (kernel.cpp:2226)
    scalar_t dd = 0.0;
    const int mc = compute_polygon_plane_coefficients(nrm, dd, v.data(), n, 1.0);
    normal[0] = nrm.x();
    normal[1] = nrm.y();
    normal[2] = nrm.z();
    *d = dd;
    return mc;
}

char ref_segment_plane_type(const double* q, const double* r, const double* verts, int n, const double* normal, int mc)
{
    return compute_segment_plane_intersection_type(vec3(q[0], q[1], q[2]), vec3(r[0], r[1], r[2]), to_vec(verts, n),
        vec3(normal[0], normal[1], normal[2]), mc, 1.0);
}

char ref_segment_plane_intersection(double* p, const double* normal, double d, const double* q, const double* r)
{
    vec3 pp(0., 0., 0.);
    const char c = compute_segment_plane_intersection(pp, vec3(normal[0], normal[1], normal[2]), d, vec3(q[0], q[1], q[2]),
        vec3(r[0], r[1], r[2]));
    p[0] = pp.x();
    p[1] = pp.y();
    p[2] = pp.z();
    return c;
}

char ref_point_in_polygon(const double* p, const double* verts, int n, const double* normal, int mc)
{
    return compute_point_in_polygon_test(vec3(p[0], p[1], p[2]), to_vec(verts, n), vec3(normal[0], normal[1], normal[2]), mc, 1.0);
}

int ref_vertex_parameters(unsigned flags, const void* src, unsigned nsv, const void* cut, unsigned ncv, double* com,
    double* shift, double* src_bbox, double* cut_bbox)
{
    double mult = 1;
    vec3_<double> t, c, smn, smx, cmn, cmx;
    const bool ok = calculate_vertex_parameters(mult, t, c, smn, smx, cmn, cmx, flags, src, nsv, cut, ncv);
    for (int j = 0; j < 3; ++j) {
        com[j] = c[j];
        shift[j] = t[j];
        src_bbox[j] = smn[j];
        src_bbox[3 + j] = smx[j];
        cut_bbox[j] = cmn[j];
        cut_bbox[3 + j] = cmx[j];
    }
    return ok ? 1 : 0;
}
}
