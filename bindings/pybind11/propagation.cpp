// bindings/pybind11/propagation.cpp -- the reference-side binding: what a GRAND+ maintainer puts in place of
// /root/reference/precompute/propagation.cpp:8-12 + class Graph (/root/reference/precompute/graph.h:17-133) to run
// GFPush on the B200 through the C ABI of libgrandplus_b200.so (include/grandplus_b200.h).
// Same module name (`from precompute import propagation`, model.py:9), same class, constructor and method signature
// (model.py:251,268; model_mag.py:271,289), same in-place fill of the caller's arrays.
//
//   g++ -O2 -shared -std=c++17 -fPIC $(python3 -m pybind11 --includes) -I<repo>/include propagation.cpp \
//       -L<repo>/grand-plus_b200 -lgrandplus_b200 -Wl,-rpath,<repo>/grand-plus_b200 \
//       -o precompute/propagation$(python3-config --extension-suffix)          (bindings/pybind11/Makefile)
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <stdexcept>
#include <string>

#include "grandplus_b200.h"

namespace py = pybind11;

namespace {

using InI32 = py::array_t<int32_t, py::array::c_style | py::array::forcecast>;   // graph.h:32: array_t<int> force-casts
using InF64 = py::array_t<double, py::array::c_style | py::array::forcecast>;

[[noreturn]] void fail(int status) {
    const char *msg = gp_last_error();
    const std::string text = std::string("libgrandplus_b200 error ") + std::to_string(status) + ": " + (msg ? msg : "");
    if (status == GP_ERR_INVALID) throw py::value_error(text);
    throw std::runtime_error(text);
}

// The reference writes through whatever pybind hands it; a wrong dtype silently fills a temporary (graph.h:63-68).
// Refuse instead: outputs must be the caller's own C-contiguous arrays of the exact dtype.
template <class T>
T *out_ptr(py::array &a, const char *name, py::ssize_t need) {
    if (!py::isinstance<py::array_t<T>>(a) || !(a.flags() & py::array::c_style) || !a.writeable())
        throw py::value_error(std::string(name) + " must be a writeable C-contiguous numpy array of the reference's dtype");
    if (a.size() < need) throw py::value_error(std::string(name) + " is shorter than len(node_idx) * K");
    return static_cast<T *>(a.mutable_data());
}

class Graph {
public:
    Graph(InI32 indptr, InI32 indices, int seed) {
        if (indptr.ndim() != 1 || indices.ndim() != 1 || indptr.size() < 2) throw py::value_error("indptr/indices must be 1-D CSR arrays");
        int device = 0;
        const int rc = gp_graph_create(indptr.data(), indptr.size() - 1, indices.data(), indices.size(), seed, device, &g_);
        if (rc != GP_OK) fail(rc);
    }
    ~Graph() { gp_graph_destroy(g_); }
    Graph(const Graph &) = delete;
    Graph &operator=(const Graph &) = delete;

    // graph.h:53: gfpush_omp(node_idx, row_idx, col_idx, value, coef, rmax, K)
    void gfpush_omp(InI32 node_idx, py::array row_idx, py::array col_idx, py::array value, InF64 coef, double rmax, int K) {
        const py::ssize_t S = node_idx.size();
        if (K < 1) throw py::value_error("K must be positive");
        int32_t *row = out_ptr<int32_t>(row_idx, "row_idx", S * K);
        int32_t *col = out_ptr<int32_t>(col_idx, "col_idx", S * K);
        double *val = out_ptr<double>(value, "value", S * K);
        int rc;
        {
            py::gil_scoped_release nogil;   // propagation.cpp:9-11 holds the GIL for the whole push
            rc = gp_gfpush(g_, node_idx.data(), S, coef.data(), (int32_t)coef.size(), rmax, K, row, col, val);
        }
        if (rc != GP_OK) fail(rc);
    }

private:
    gp_graph *g_ = nullptr;
};

}  // namespace

PYBIND11_MODULE(propagation, m) {   // same module and names as propagation.cpp:8-12
    m.doc() = "GRAND+ GFPush on B200 through libgrandplus_b200.so (drop-in for precompute.propagation)";
    py::class_<Graph>(m, "Graph")
        .def(py::init<InI32, InI32, int>())
        .def("gfpush_omp", &Graph::gfpush_omp);
}
