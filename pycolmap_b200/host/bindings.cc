// bindings.cc -- pybind11 module `pycolmap_b200._core`: the C++ host side of the drop-in boundary.
//
// Mirrors, name for name and keyword for keyword, the part of the reference's Python surface that
// reaches the matching / verification hot path:
//   match_exhaustive, match_sequential, verify_matches        R:pipeline/match_features.h:22-68, 219-260
//   SiftMatchingOptions, Exhaustive/SequentialMatchingOptions R:pipeline/match_features.h:71-152
//   TwoViewGeometryOptions, TwoViewGeometry(Configuration)    R:estimators/two_view_geometry.h:41-93
//   estimate_(calibrated_)two_view_geometry, squared_sampson_error   R:estimators/two_view_geometry.h:95-175
//   essential / fundamental / homography_matrix_estimation    R:estimators/essential_matrix.h:19-103,
//                                                             fundamental_matrix.h:17-50, homography_matrix.h:17-48
//   RANSACOptions                                             R:optim/bindings.h:7-27
//   Device, has_cuda                                          R:utils.h:9-31, R:main.cc:98-106
//   Database (open/close/counters/read_two_view_geometry/image_pair_to_pair_id)  R:scene/database.h:10-47
// Everything below the option structs goes through the C ABI of include/b200match.h; there is no CPU path.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <array>
#include <atomic>
#include <chrono>
#include <cstring>
#include <functional>
#include <thread>

#include "../csrc/camera_models.h"
#include "controllers.h"
#include "dataclass.h"
#include "database.h"

namespace py = pybind11;
using namespace b2mh;
using namespace pybind11::literals;

namespace {

enum class Device { AUTO = -1, CPU = 0, CUDA = 1 };  // R:utils.h:9

enum class TwoViewGeometryConfiguration {  // R:estimators/two_view_geometry.h:67-80
  UNDEFINED = 0, DEGENERATE, CALIBRATED, UNCALIBRATED, PLANAR, PANORAMIC, PLANAR_OR_PANORAMIC, WATERMARK, MULTIPLE
};

using ArrD = py::array_t<double, py::array::c_style | py::array::forcecast>;
using ArrU8 = py::array_t<uint8_t, py::array::c_style | py::array::forcecast>;
using ArrF32 = py::array_t<float, py::array::c_style | py::array::forcecast>;
using ArrU32 = py::array_t<uint32_t, py::array::c_style | py::array::forcecast>;
using ArrI32 = py::array_t<int32_t, py::array::c_style | py::array::forcecast>;

py::array_t<double> MatToNumpy(const double* m) {
  py::array_t<double> a(std::vector<py::ssize_t>{3, 3});
  memcpy(a.mutable_data(), m, 72);
  return a;
}

py::array_t<uint32_t> MatchesToNumpy(const uint32_t* m, int64_t n) {
  py::array_t<uint32_t> a(std::vector<py::ssize_t>{static_cast<py::ssize_t>(n), 2});
  if (n > 0) memcpy(a.mutable_data(), m, static_cast<size_t>(n) * 8);
  return a;
}

std::string FsPath(const py::object& p) { return py::module_::import("os").attr("fspath")(p).cast<std::string>(); }

void CheckPoints(const ArrD& p, const char* name) {
  if (p.ndim() != 2 || p.shape(1) != 2)
    throw std::invalid_argument(std::string("[bindings.cc] Check Failed: ") + name + " is an N x 2 array");
}

// A camera is a dict {model, width, height, params, has_prior_focal_length} or any object with those
// attributes (a pycolmap.Camera: `model` may be an enum with a .name, R:scene/camera.h:20-213).
b2m_camera CameraFromPython(const py::object& cam) {
  auto get = [&](const char* key, py::object def) -> py::object {
    if (py::isinstance<py::dict>(cam)) {
      const py::dict d = cam;
      return d.contains(key) ? py::object(d[key]) : def;
    }
    return py::hasattr(cam, key) ? py::object(cam.attr(key)) : def;
  };
  py::object model = get("model", py::none());
  if (model.is_none()) model = get("model_id", py::int_(0));
  if (py::hasattr(model, "name")) model = model.attr("name");
  CameraRow row;
  if (py::isinstance<py::str>(model)) {
    const std::string name = model.cast<std::string>();
    static const char* kNames[] = {"SIMPLE_PINHOLE", "PINHOLE", "SIMPLE_RADIAL", "RADIAL", "OPENCV", "OPENCV_FISHEYE",
                                   "FULL_OPENCV", "FOV", "SIMPLE_RADIAL_FISHEYE", "RADIAL_FISHEYE", "THIN_PRISM_FISHEYE"};
    row.model = -1;
    for (int i = 0; i < 11; ++i)
      if (name == kNames[i]) row.model = i;
    if (row.model < 0) throw std::invalid_argument("[bindings.cc] unknown camera model " + name);
  } else {
    row.model = model.cast<int>();
  }
  row.width = get("width", py::int_(0)).cast<int64_t>();
  row.height = get("height", py::int_(0)).cast<int64_t>();
  row.params = get("params", py::list()).cast<std::vector<double>>();
  row.has_prior_focal_length = py::bool_(get("has_prior_focal_length", py::bool_(false)));
  return ToAbi(row);
}

py::dict CameraToDict(const CameraRow& c) {
  return py::dict("model"_a = c.model, "width"_a = c.width, "height"_a = c.height, "params"_a = c.params,
                  "has_prior_focal_length"_a = c.has_prior_focal_length ? 1 : 0);
}

// ---- TwoViewGeometry (R:estimators/two_view_geometry.h:82-93) --------------------------------------
// Minimal stand-ins for pycolmap.Rotation3d / Rigid3d (R:geometry bindings): x_cam2 = rotation * x_cam1 + translation.
struct Rotation3d {
  std::array<double, 4> wxyz{1.0, 0.0, 0.0, 0.0};
};
struct Rigid3d {
  Rotation3d rotation;
  std::array<double, 3> translation{0.0, 0.0, 0.0};
};

struct TwoViewGeometry {
  TwoViewGeometryConfiguration config = TwoViewGeometryConfiguration::UNDEFINED;
  Mat3 E{}, F{}, H{};
  std::vector<uint32_t> inlier_matches;
  double tri_angle = 0.0;
  std::array<double, 4> qvec{1.0, 0.0, 0.0, 0.0};  // cam2_from_cam1 (w, x, y, z); identity unless a pose was recovered
  std::array<double, 3> tvec{0.0, 0.0, 0.0};
  int nE = 0, nF = 0, nH = 0;  // diagnostic inlier counts of the three LO-RANSAC runs

  void Invert() {
    InvertPose(&qvec, &tvec);
    F = Transposed(F);
    E = Transposed(E);
    Mat3 inv;
    bool zero = true;
    for (double v : H) zero = zero && v == 0.0;
    if (!zero && Invert3x3(H, &inv)) H = inv;
    for (size_t i = 0; i + 1 < inlier_matches.size(); i += 2) std::swap(inlier_matches[i], inlier_matches[i + 1]);
  }
};

// ---- PyWait (R:helpers.h:335-347): run `fn` on a worker thread with the GIL released; the calling
// thread polls for signals so that Ctrl-C stops the GPU work between batches and surfaces as
// KeyboardInterrupt.
template <typename Fn>
void RunInterruptible(Fn&& fn) {
  std::exception_ptr error;
  PyObject *et = nullptr, *ev = nullptr, *tb = nullptr;
  bool interrupted = false;
  {
    py::gil_scoped_release release;
    std::atomic<bool> done{false};
    std::thread worker([&] {
      try {
        fn();
      } catch (...) {
        error = std::current_exception();
      }
      done = true;
    });
    while (!done) {
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
      if (interrupted) continue;
      py::gil_scoped_acquire acquire;
      if (PyErr_CheckSignals() != 0) {
        PyErr_Fetch(&et, &ev, &tb);
        interrupted = true;
        Engine::RequestStopAll();
      }
    }
    worker.join();
  }
  if (interrupted) {
    PyErr_Restore(et, ev, tb);
    throw py::error_already_set();
  }
  if (error) std::rethrow_exception(error);
}

std::vector<int> ResolveDevices(Device device, const SiftMatchingOptions& sift) {
  // IsGPU / VerifyGPUParams (R:utils.h:11-31): `auto` is the GPU; this build has nothing else.
  if (device == Device::CPU)
    throw std::invalid_argument("[bindings.cc] pycolmap_b200 has no CPU path: use Device.auto or Device.cuda");
  return ParseGpuIndices(sift.gpu_index);
}

// ---- estimators --------------------------------------------------------------------------------------
// `get_ctx` is called only after every argument check passed (checks come before any GPU work).
TwoViewGeometry EstimateTvg(const std::function<b2m_ctx*()>& get_ctx, const py::object& camera1, const ArrD& points1, const py::object& camera2,
                            const ArrD& points2, const py::object& matches, const TwoViewGeometryOptions& options) {
  CheckPoints(points1, "points1");
  CheckPoints(points2, "points2");
  const b2m_camera c1 = CameraFromPython(camera1), c2 = CameraFromPython(camera2);
  const b2m_tvg_opts opts = ToAbi(options);
  ArrU32 m;
  const uint32_t* mptr = nullptr;
  int64_t n_m = points1.shape(0);
  if (!matches.is_none()) {
    m = ArrU32::ensure(matches);
    if (!m || m.ndim() != 2 || m.shape(1) != 2)
      throw std::invalid_argument("[bindings.cc] Check Failed: matches is an N x 2 uint32 array");
    mptr = m.data();
    n_m = m.shape(0);
  } else if (points1.shape(0) != points2.shape(0)) {
    throw std::invalid_argument("[two_view_geometry.h:137] Check Failed: points1.size() == points2.size()");
  }
  b2m_tvg_result r;
  memset(&r, 0, sizeof(r));
  r.struct_size = sizeof(r);
  std::vector<uint32_t> inl(static_cast<size_t>(std::max<int64_t>(1, n_m)) * 2);
  b2m_ctx* ctx = get_ctx();
  int rc;
  {
    py::gil_scoped_release release;
    rc = b2m_estimate_two_view_geometry(ctx, &c1, points1.data(), points1.shape(0), &c2, points2.data(),
                                        points2.shape(0), mptr, n_m, &opts, &r, inl.data());
  }
  ThrowOnError(ctx, rc);
  TwoViewGeometry g;
  g.config = static_cast<TwoViewGeometryConfiguration>(r.config);
  std::copy(r.E, r.E + 9, g.E.begin());
  std::copy(r.F, r.F + 9, g.F.begin());
  std::copy(r.H, r.H + 9, g.H.begin());
  inl.resize(static_cast<size_t>(r.n_inliers) * 2);
  g.inlier_matches = std::move(inl);
  g.nE = r.nE; g.nF = r.nF; g.nH = r.nH;
  std::copy(r.qvec, r.qvec + 4, g.qvec.begin());
  std::copy(r.tvec, r.tvec + 3, g.tvec.begin());
  g.tri_angle = r.tri_angle;
  return g;
}

// Single-model LO-RANSAC; returns None on failure like the reference (R:estimators/fundamental_matrix.h:31-33).
py::object RansacModel(b2m_ctx* ctx, int kind, const char* key, const ArrD& p1, const ArrD& p2,
                       const RANSACOptions& options) {
  const b2m_ransac_opts o = ToAbi(options);
  const int64_t m = p1.shape(0);
  double model[9] = {0};
  std::vector<uint8_t> mask(static_cast<size_t>(std::max<int64_t>(1, m)));
  int64_t n_inl = 0;
  int32_t ok = 0;
  int rc;
  {
    py::gil_scoped_release release;
    rc = b2m_ransac_model(ctx, kind, p1.data(), p2.data(), m, &o, model, mask.data(), &n_inl, &ok);
  }
  ThrowOnError(ctx, rc);
  if (!ok) return py::none();
  py::array_t<bool> inliers(static_cast<py::ssize_t>(m));
  bool* ip = inliers.mutable_data();
  for (int64_t i = 0; i < m; ++i) ip[i] = mask[i] != 0;
  py::dict out;
  out[key] = MatToNumpy(model);
  out["num_inliers"] = n_inl;
  out["inliers"] = inliers;
  return out;
}

void CheckSameLength(const ArrD& p1, const ArrD& p2, const char* where) {
  CheckPoints(p1, "points1");
  CheckPoints(p2, "points2");
  if (p1.shape(0) != p2.shape(0))
    throw std::invalid_argument(std::string("[") + where + "] Check Failed: points1.size() == points2.size()");
}

// ---- low-level context (what tests and benchmarks drive directly) ------------------------------------
struct CoreResults {
  b2m_results* r = nullptr;
  ~CoreResults() { Free(); }
  void Free() {
    if (r) b2m_results_free(r);
    r = nullptr;
  }
  b2m_pair_view View(int64_t k) const {
    if (!r) throw std::runtime_error("[bindings.cc] results were freed");
    b2m_pair_view v;
    memset(&v, 0, sizeof(v));
    v.struct_size = sizeof(v);
    if (b2m_results_get(r, k, &v) != B2M_OK) throw py::index_error(std::to_string(k));
    return v;
  }
};

struct CoreContext {
  b2m_ctx* ctx = nullptr;
  CoreContext(int device, uint64_t seed, int pair_batch) {
    b2m_device_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.device = device;
    cfg.seed = seed;
    cfg.pair_batch = pair_batch;
    ThrowOnError(nullptr, b2m_create(&cfg, &ctx));
  }
  ~CoreContext() { Close(); }
  void Close() {
    if (ctx) b2m_destroy(ctx);
    ctx = nullptr;
  }
  b2m_ctx* Handle() const {
    if (!ctx) throw std::runtime_error("[bindings.cc] context is closed");
    return ctx;
  }
};

std::vector<b2m_camera> CamerasFromPython(const py::object& cams) {
  std::vector<b2m_camera> out;
  for (const auto& c : cams) out.push_back(CameraFromPython(py::reinterpret_borrow<py::object>(c)));
  return out;
}

}  // namespace

PYBIND11_MODULE(_core, m) {
  m.doc() = "C++ host layer of pycolmap_b200 (pybind11 over the C ABI of libb200match.so)";
  m.attr("has_cuda") = true;  // R:main.cc:98
  m.attr("abi_version") = b2m_abi_version();

  py::register_exception_translator([](std::exception_ptr p) {
    try {
      if (p) std::rethrow_exception(p);
    } catch (const StoppedError&) {
      PyErr_SetString(PyExc_KeyboardInterrupt, "stopped");
    }
  });

  // ---- enums ----
  auto device = py::enum_<Device>(m, "Device").value("auto", Device::AUTO).value("cpu", Device::CPU).value("cuda", Device::CUDA);
  AddStringConstructor(device);
  auto config = py::enum_<TwoViewGeometryConfiguration>(m, "TwoViewGeometryConfiguration")
                    .value("UNDEFINED", TwoViewGeometryConfiguration::UNDEFINED)
                    .value("DEGENERATE", TwoViewGeometryConfiguration::DEGENERATE)
                    .value("CALIBRATED", TwoViewGeometryConfiguration::CALIBRATED)
                    .value("UNCALIBRATED", TwoViewGeometryConfiguration::UNCALIBRATED)
                    .value("PLANAR", TwoViewGeometryConfiguration::PLANAR)
                    .value("PANORAMIC", TwoViewGeometryConfiguration::PANORAMIC)
                    .value("PLANAR_OR_PANORAMIC", TwoViewGeometryConfiguration::PLANAR_OR_PANORAMIC)
                    .value("WATERMARK", TwoViewGeometryConfiguration::WATERMARK)
                    .value("MULTIPLE", TwoViewGeometryConfiguration::MULTIPLE);
  AddStringConstructor(config);

  // ---- option classes ----
  OptionsClass<RANSACOptions>(m, "RANSACOptions")
      .field("max_error", &RANSACOptions::max_error)
      .field("min_inlier_ratio", &RANSACOptions::min_inlier_ratio)
      .field("confidence", &RANSACOptions::confidence)
      .field("dyn_num_trials_multiplier", &RANSACOptions::dyn_num_trials_multiplier)
      .field("min_num_trials", &RANSACOptions::min_num_trials)
      .field("max_num_trials", &RANSACOptions::max_num_trials)
      .Finish();
  OptionsClass<SiftMatchingOptions>(m, "SiftMatchingOptions")
      .field("num_threads", &SiftMatchingOptions::num_threads)
      .field("gpu_index", &SiftMatchingOptions::gpu_index,
             "Index of the GPU used for feature matching. For multi-GPU matching, you should separate multiple "
             "GPU indices by comma, e.g., \"0,1,2,3\".")
      .field("max_ratio", &SiftMatchingOptions::max_ratio, "Maximum distance ratio between first and second best match.")
      .field("max_distance", &SiftMatchingOptions::max_distance, "Maximum distance to best match.")
      .field("cross_check", &SiftMatchingOptions::cross_check, "Whether to enable cross checking in matching.")
      .field("max_num_matches", &SiftMatchingOptions::max_num_matches, "Maximum number of matches.")
      .field("guided_matching", &SiftMatchingOptions::guided_matching,
             "Whether to perform guided matching, if geometric verification succeeds.")
      .Finish();
  OptionsClass<ExhaustiveMatchingOptions>(m, "ExhaustiveMatchingOptions")
      .field("block_size", &ExhaustiveMatchingOptions::block_size)
      .Finish();
  OptionsClass<SequentialMatchingOptions>(m, "SequentialMatchingOptions")
      .field("overlap", &SequentialMatchingOptions::overlap, "Number of overlapping image pairs.")
      .field("quadratic_overlap", &SequentialMatchingOptions::quadratic_overlap,
             "Whether to match images against their quadratic neighbors.")
      .field("loop_detection", &SequentialMatchingOptions::loop_detection,
             "Loop detection is invoked every `loop_detection_period` images.")
      .field("loop_detection_period", &SequentialMatchingOptions::loop_detection_period)
      .field("loop_detection_num_images", &SequentialMatchingOptions::loop_detection_num_images)
      .field("loop_detection_num_nearest_neighbors", &SequentialMatchingOptions::loop_detection_num_nearest_neighbors)
      .field("loop_detection_num_checks", &SequentialMatchingOptions::loop_detection_num_checks)
      .field("loop_detection_num_images_after_verification",
             &SequentialMatchingOptions::loop_detection_num_images_after_verification)
      .field("loop_detection_max_num_features", &SequentialMatchingOptions::loop_detection_max_num_features)
      .field("vocab_tree_path", &SequentialMatchingOptions::vocab_tree_path)
      .Finish();
  OptionsClass<SpatialMatchingOptions>(m, "SpatialMatchingOptions")
      .field("is_gps", &SpatialMatchingOptions::is_gps,
             "Whether the location priors in the database are GPS coordinates in the form of longitude and latitude "
             "coordinates in degrees.")
      .field("ignore_z", &SpatialMatchingOptions::ignore_z, "Whether to ignore the Z-component of the location prior.")
      .field("max_num_neighbors", &SpatialMatchingOptions::max_num_neighbors, "The maximum number of nearest neighbors to match.")
      .field("max_distance", &SpatialMatchingOptions::max_distance,
             "The maximum distance between the query and nearest neighbor [meters].")
      .Finish();
  OptionsClass<TwoViewGeometryOptions>(m, "TwoViewGeometryOptions")
      .field("min_num_inliers", &TwoViewGeometryOptions::min_num_inliers)
      .field("min_E_F_inlier_ratio", &TwoViewGeometryOptions::min_E_F_inlier_ratio)
      .field("max_H_inlier_ratio", &TwoViewGeometryOptions::max_H_inlier_ratio)
      .field("watermark_min_inlier_ratio", &TwoViewGeometryOptions::watermark_min_inlier_ratio)
      .field("watermark_border_size", &TwoViewGeometryOptions::watermark_border_size)
      .field("detect_watermark", &TwoViewGeometryOptions::detect_watermark)
      .field("multiple_ignore_watermark", &TwoViewGeometryOptions::multiple_ignore_watermark)
      .field("force_H_use", &TwoViewGeometryOptions::force_H_use)
      .field("compute_relative_pose", &TwoViewGeometryOptions::compute_relative_pose)
      .field("multiple_models", &TwoViewGeometryOptions::multiple_models)
      .field("ransac", &TwoViewGeometryOptions::ransac)
      .Finish();

  py::class_<Rotation3d>(m, "Rotation3d")
      .def(py::init<>())
      .def(py::init([](const std::array<double, 4>& xyzw) {   // Eigen coefficient order, like the reference
             Rotation3d r;
             r.wxyz = {xyzw[3], xyzw[0], xyzw[1], xyzw[2]};
             return r;
           }),
           "xyzw"_a)
      .def_property_readonly("quat", [](const Rotation3d& r) {   // (x, y, z, w) like the reference's Eigen coefficients
        py::array_t<double> a(4);
        double* p = a.mutable_data();
        p[0] = r.wxyz[1]; p[1] = r.wxyz[2]; p[2] = r.wxyz[3]; p[3] = r.wxyz[0];
        return a;
      })
      .def("matrix", [](const Rotation3d& r) { return MatToNumpy(QuatToRotation(r.wxyz).data()); });
  py::class_<Rigid3d>(m, "Rigid3d")
      .def(py::init<>())
      .def(py::init([](const Rotation3d& rotation, const std::array<double, 3>& translation) {
             return Rigid3d{rotation, translation};
           }),
           "rotation"_a, "translation"_a)
      .def_readonly("rotation", &Rigid3d::rotation)
      .def_property_readonly("translation", [](const Rigid3d& g) {
        py::array_t<double> a(3);
        std::copy(g.translation.begin(), g.translation.end(), a.mutable_data());
        return a;
      })
      .def("matrix", [](const Rigid3d& g) {
        const Mat3 R = QuatToRotation(g.rotation.wxyz);
        py::array_t<double> a(std::vector<py::ssize_t>{3, 4});
        double* p = a.mutable_data();
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) p[i * 4 + j] = R[i * 3 + j];
          p[i * 4 + 3] = g.translation[i];
        }
        return a;
      })
      .def("inverse", [](const Rigid3d& g) {
        Rigid3d o = g;
        InvertPose(&o.rotation.wxyz, &o.translation);
        return o;
      });

  // ---- TwoViewGeometry ----
  py::class_<TwoViewGeometry>(m, "TwoViewGeometry")
      .def(py::init<>())
      .def(py::init([](TwoViewGeometryConfiguration config, const py::object& E, const py::object& F,
                       const py::object& H, const py::object& inlier_matches, double tri_angle, const py::object& qvec,
                       const py::object& tvec) {
             TwoViewGeometry g;
             g.config = config;
             auto mat = [](const py::object& o, Mat3* out) {
               if (o.is_none()) return;
               const ArrD a = ArrD::ensure(o);
               if (!a || a.size() != 9) throw std::invalid_argument("[bindings.cc] Check Failed: matrix is 3 x 3");
               std::copy(a.data(), a.data() + 9, out->begin());
             };
             mat(E, &g.E); mat(F, &g.F); mat(H, &g.H);
             if (!inlier_matches.is_none()) {
               const ArrU32 a = ArrU32::ensure(inlier_matches);
               if (!a || a.size() % 2) throw std::invalid_argument("[bindings.cc] Check Failed: matches is N x 2");
               g.inlier_matches.assign(a.data(), a.data() + a.size());
             }
             g.tri_angle = tri_angle;
             auto vec = [](const py::object& o, double* out, py::ssize_t n, const char* what) {
               if (o.is_none()) return;
               const ArrD a = ArrD::ensure(o);
               if (!a || a.size() != n) throw std::invalid_argument(std::string("[bindings.cc] Check Failed: ") + what);
               std::copy(a.data(), a.data() + n, out);
             };
             vec(qvec, g.qvec.data(), 4, "qvec = (w, x, y, z)");
             vec(tvec, g.tvec.data(), 3, "tvec has 3 entries");
             return g;
           }),
           "config"_a = TwoViewGeometryConfiguration::UNDEFINED, "E"_a = py::none(), "F"_a = py::none(),
           "H"_a = py::none(), "inlier_matches"_a = py::none(), "tri_angle"_a = 0.0, "qvec"_a = py::none(),
           "tvec"_a = py::none())
      .def_property_readonly("config", [](const TwoViewGeometry& g) { return g.config; })
      .def_property_readonly("E", [](const TwoViewGeometry& g) { return MatToNumpy(g.E.data()); })
      .def_property_readonly("F", [](const TwoViewGeometry& g) { return MatToNumpy(g.F.data()); })
      .def_property_readonly("H", [](const TwoViewGeometry& g) { return MatToNumpy(g.H.data()); })
      .def_property_readonly("cam2_from_cam1",
                             [](const TwoViewGeometry& g) { return Rigid3d{Rotation3d{g.qvec}, g.tvec}; },
                             "Relative pose (identity unless options.compute_relative_pose recovered one).")
      .def_property_readonly("inlier_matches",
                             [](const TwoViewGeometry& g) {
                               return MatchesToNumpy(g.inlier_matches.data(),
                                                     static_cast<int64_t>(g.inlier_matches.size() / 2));
                             })
      .def_readonly("tri_angle", &TwoViewGeometry::tri_angle)
      .def_property_readonly("num_inliers_EFH", [](const TwoViewGeometry& g) { return py::make_tuple(g.nE, g.nF, g.nH); })
      .def("invert", &TwoViewGeometry::Invert)
      .def("__repr__", [](const TwoViewGeometry& g) {
        return "TwoViewGeometry(config=" + py::str(py::cast(g.config).attr("name")).cast<std::string>() +
               ", num_inliers=" + std::to_string(g.inlier_matches.size() / 2) + ")";
      });

  // ---- pipelines ----
  m.def("match_exhaustive",
        [](const py::object& database_path, SiftMatchingOptions sift_options, ExhaustiveMatchingOptions matching_options,
           TwoViewGeometryOptions verification_options, Device device) {
          const std::string path = FsPath(database_path);
          CheckFileExists(path, "match_features.h:32");  // before the device check, R:match_features.h:32-38
          const std::vector<int> dev = ResolveDevices(device, sift_options);
          RunInterruptible([&] { MatchExhaustive(path, sift_options, matching_options, verification_options, dev); });
        },
        "database_path"_a, "sift_options"_a = SiftMatchingOptions(), "matching_options"_a = ExhaustiveMatchingOptions(),
        "verification_options"_a = TwoViewGeometryOptions(), "device"_a = Device::AUTO,
        "Exhaustive feature matching + geometric verification of every image pair of the database");
  m.def("match_sequential",
        [](const py::object& database_path, SiftMatchingOptions sift_options, SequentialMatchingOptions matching_options,
           TwoViewGeometryOptions verification_options, Device device) {
          const std::string path = FsPath(database_path);
          CheckFileExists(path, "match_features.h:32");  // before the device check, R:match_features.h:32-38
          const std::vector<int> dev = ResolveDevices(device, sift_options);
          RunInterruptible([&] { MatchSequential(path, sift_options, matching_options, verification_options, dev); });
        },
        "database_path"_a, "sift_options"_a = SiftMatchingOptions(), "matching_options"_a = SequentialMatchingOptions(),
        "verification_options"_a = TwoViewGeometryOptions(), "device"_a = Device::AUTO,
        "Sequential feature matching (images ordered by name; overlap / quadratic overlap)");
  m.def("match_spatial",
        [](const py::object& database_path, SiftMatchingOptions sift_options, SpatialMatchingOptions matching_options,
           TwoViewGeometryOptions verification_options, Device device) {
          const std::string path = FsPath(database_path);
          CheckFileExists(path, "match_features.h:32");
          const std::vector<int> dev = ResolveDevices(device, sift_options);
          RunInterruptible([&] { MatchSpatial(path, sift_options, matching_options, verification_options, dev); });
        },
        "database_path"_a, "sift_options"_a = SiftMatchingOptions(), "matching_options"_a = SpatialMatchingOptions(),
        "verification_options"_a = TwoViewGeometryOptions(), "device"_a = Device::AUTO,
        "Spatial feature matching (nearest neighbours by location prior)");
  m.def("match_vocabtree",
        [](const py::object&, const py::args&, const py::kwargs&) {
          throw std::invalid_argument("[bindings.cc] match_vocabtree needs a vocabulary tree index: out of scope (SURVEY.md section 8(f) item 4)");
        },
        "database_path"_a, "Not available: vocabulary-tree retrieval is outside the hot path this library covers.");
  m.def("verify_matches",
        [](const py::object& database_path, const py::object& pairs_path, TwoViewGeometryOptions options) {
          const std::string db = FsPath(database_path), pairs = FsPath(pairs_path);
          RunInterruptible([&] { VerifyMatches(db, pairs, options); });
        },
        "database_path"_a, "pairs_path"_a, "options"_a = TwoViewGeometryOptions(),
        "Run geometric verification of the matches of the listed pairs");

  // ---- estimators ----
  m.def("estimate_two_view_geometry",
        [](const py::object& camera1, const ArrD& points1, const py::object& camera2, const ArrD& points2,
           const py::object& matches, const TwoViewGeometryOptions& options) {
          return EstimateTvg([] { return Engine::Get(0); }, camera1, points1, camera2, points2, matches, options);
        },
        "camera1"_a, "points1"_a, "camera2"_a, "points2"_a, "matches"_a = py::none(),
        "options"_a = TwoViewGeometryOptions());
  m.def("estimate_calibrated_two_view_geometry",
        [](const py::object& camera1, const ArrD& points1, const py::object& camera2, const ArrD& points2,
           const py::object& matches, const TwoViewGeometryOptions& options) {
          // EstimateCalibratedTwoViewGeometry: the E branch runs whatever the prior flags say
          auto forced = [](const py::object& cam) {
            b2m_camera c = CameraFromPython(cam);
            return py::dict("model"_a = c.model, "width"_a = c.width, "height"_a = c.height,
                            "params"_a = std::vector<double>(c.params, c.params + b2m::cam::num_params(c.model)),
                            "has_prior_focal_length"_a = 1);
          };
          return EstimateTvg([] { return Engine::Get(0); }, forced(camera1), points1, forced(camera2), points2, matches, options);
        },
        "camera1"_a, "points1"_a, "camera2"_a, "points2"_a, "matches"_a = py::none(),
        "options"_a = TwoViewGeometryOptions());
  m.def("estimate_two_view_geometries",
        [](const py::list& problems, const TwoViewGeometryOptions& options) {
          // Batched estimate_two_view_geometry: every problem is a tuple
          // (camera1, points1, camera2, points2[, matches]); one GPU launch verifies them all.
          const size_t n = problems.size();
          std::vector<ArrD> p1(n), p2(n);
          std::vector<ArrU32> mm(n);
          std::vector<b2m_tvg_problem> q(n);
          std::vector<std::vector<uint32_t>> inl(n);
          std::vector<uint32_t*> inl_ptr(n);
          for (size_t k = 0; k < n; ++k) {
            const py::tuple t = py::tuple(problems[k]);
            if (t.size() != 4 && t.size() != 5)
              throw std::invalid_argument("[bindings.cc] Check Failed: problem = (camera1, points1, camera2, points2[, matches])");
            p1[k] = ArrD::ensure(t[1]);
            p2[k] = ArrD::ensure(t[3]);
            if (!p1[k] || !p2[k]) throw std::invalid_argument("[bindings.cc] Check Failed: points are N x 2 float64 arrays");
            CheckPoints(p1[k], "points1");
            CheckPoints(p2[k], "points2");
            memset(&q[k], 0, sizeof(q[k]));
            q[k].struct_size = sizeof(q[k]);
            q[k].cam1 = CameraFromPython(py::reinterpret_borrow<py::object>(t[0]));
            q[k].cam2 = CameraFromPython(py::reinterpret_borrow<py::object>(t[2]));
            q[k].points1 = p1[k].data();
            q[k].n1 = p1[k].shape(0);
            q[k].points2 = p2[k].data();
            q[k].n2 = p2[k].shape(0);
            int64_t m = q[k].n1;
            if (t.size() == 5 && !t[4].is_none()) {
              mm[k] = ArrU32::ensure(t[4]);
              if (!mm[k] || mm[k].ndim() != 2 || mm[k].shape(1) != 2)
                throw std::invalid_argument("[bindings.cc] Check Failed: matches is an N x 2 uint32 array");
              q[k].matches = mm[k].data();
              q[k].m = m = mm[k].shape(0);
            } else if (q[k].n1 != q[k].n2) {
              throw std::invalid_argument("[two_view_geometry.h:137] Check Failed: points1.size() == points2.size()");
            }
            inl[k].resize(static_cast<size_t>(std::max<int64_t>(1, m)) * 2);
            inl_ptr[k] = inl[k].data();
          }
          const b2m_tvg_opts opts = ToAbi(options);
          std::vector<b2m_tvg_result> r(n);
          b2m_ctx* ctx = Engine::Get(0);
          int rc;
          {
            py::gil_scoped_release release;
            rc = b2m_estimate_two_view_geometry_batch(ctx, q.data(), static_cast<int64_t>(n), &opts, r.data(), inl_ptr.data());
          }
          ThrowOnError(ctx, rc);
          std::vector<TwoViewGeometry> out(n);
          for (size_t k = 0; k < n; ++k) {
            out[k].config = static_cast<TwoViewGeometryConfiguration>(r[k].config);
            std::copy(r[k].E, r[k].E + 9, out[k].E.begin());
            std::copy(r[k].F, r[k].F + 9, out[k].F.begin());
            std::copy(r[k].H, r[k].H + 9, out[k].H.begin());
            inl[k].resize(static_cast<size_t>(r[k].n_inliers) * 2);
            out[k].inlier_matches = std::move(inl[k]);
            out[k].nE = r[k].nE; out[k].nF = r[k].nF; out[k].nH = r[k].nH;
            std::copy(r[k].qvec, r[k].qvec + 4, out[k].qvec.begin());
            std::copy(r[k].tvec, r[k].tvec + 3, out[k].tvec.begin());
            out[k].tri_angle = r[k].tri_angle;
          }
          return out;
        },
        "problems"_a, "options"_a = TwoViewGeometryOptions(),
        "Batched estimate_two_view_geometry: problems = [(camera1, points1, camera2, points2[, matches]), ...]");
  m.def("estimate_two_view_geometry_pose",
        [](const py::object& camera1, const ArrD& points1, const py::object& camera2, const ArrD& points2, TwoViewGeometry& geometry) {
          CheckPoints(points1, "points1");
          CheckPoints(points2, "points2");
          const b2m_camera c1 = CameraFromPython(camera1), c2 = CameraFromPython(camera2);
          b2m_tvg_result g;
          memset(&g, 0, sizeof(g));
          g.struct_size = sizeof(g);
          g.config = static_cast<int>(geometry.config);
          std::copy(geometry.E.begin(), geometry.E.end(), g.E);
          std::copy(geometry.H.begin(), geometry.H.end(), g.H);
          b2m_ctx* ctx = Engine::Get(0);
          ThrowOnError(ctx, b2m_estimate_two_view_geometry_pose(ctx, &c1, points1.data(), points1.shape(0), &c2, points2.data(),
                                                               points2.shape(0), geometry.inlier_matches.data(),
                                                               static_cast<int64_t>(geometry.inlier_matches.size() / 2), &g));
          if (!g.pose_valid) return false;
          geometry.config = static_cast<TwoViewGeometryConfiguration>(g.config);
          std::copy(g.qvec, g.qvec + 4, geometry.qvec.begin());
          std::copy(g.tvec, g.tvec + 3, geometry.tvec.begin());
          geometry.tri_angle = g.tri_angle;
          return true;
        },
        "camera1"_a, "points1"_a, "camera2"_a, "points2"_a, "geometry"_a,
        "Relative pose of an estimated geometry, in place (R:estimators/two_view_geometry.h:153-158).");
  m.def("fundamental_matrix_estimation",
        [](const ArrD& points1, const ArrD& points2, const RANSACOptions& estimation_options) {
          CheckSameLength(points1, points2, "fundamental_matrix.h:22");
          return RansacModel(Engine::Get(0), 1, "F", points1, points2, estimation_options);
        },
        "points2D1"_a, "points2D2"_a, "estimation_options"_a = RANSACOptions());
  m.def("homography_matrix_estimation",
        [](const ArrD& points1, const ArrD& points2, const RANSACOptions& estimation_options) {
          CheckSameLength(points1, points2, "homography_matrix.h:21");
          return RansacModel(Engine::Get(0), 2, "H", points1, points2, estimation_options);
        },
        "points2D1"_a, "points2D2"_a, "estimation_options"_a = RANSACOptions());
  m.def("essential_matrix_estimation",
        [](const ArrD& points1, const ArrD& points2, const py::object& camera1, const py::object& camera2,
           const RANSACOptions& estimation_options) {
          CheckSameLength(points1, points2, "essential_matrix.h:26");
          const b2m_camera c1 = CameraFromPython(camera1), c2 = CameraFromPython(camera2);
          // Camera::CamFromImg on both point lists (R:estimators/essential_matrix.h:31-39; on the GPU, every
          // supported model) and the threshold averaged over both cameras (:42-46).
          b2m_ctx* ctx = Engine::Get(0);
          auto normalise = [ctx](const b2m_camera& c, const ArrD& p) {
            ArrD out(std::vector<py::ssize_t>{p.shape(0), 2});
            ThrowOnError(ctx, b2m_cam_from_img(ctx, &c, p.data(), p.shape(0), out.mutable_data()));
            return out;
          };
          auto mean_f = [](const b2m_camera& c) { return b2m::cam::mean_focal_length(c.model, c.params); };
          RANSACOptions o = estimation_options;
          o.max_error = 0.5 * (o.max_error / mean_f(c1) + o.max_error / mean_f(c2));
          py::object res = RansacModel(ctx, 0, "E", normalise(c1, points1), normalise(c2, points2), o);
          if (res.is_none()) return res;
          // PoseFromEssentialMatrix on the inliers (R:estimators/essential_matrix.h:62-83), on the GPU
          py::dict d = res;
          const py::array_t<bool> mask = py::array_t<bool>::ensure(d["inliers"]);
          std::vector<uint32_t> inl;
          for (py::ssize_t i = 0; i < mask.size(); ++i)
            if (mask.data()[i]) {
              inl.push_back(static_cast<uint32_t>(i));
              inl.push_back(static_cast<uint32_t>(i));
            }
          b2m_tvg_result g;
          memset(&g, 0, sizeof(g));
          g.struct_size = sizeof(g);
          g.config = B2M_CALIBRATED;
          const ArrD E = ArrD::ensure(py::object(d["E"]));
          std::copy(E.data(), E.data() + 9, g.E);
          ThrowOnError(ctx, b2m_estimate_two_view_geometry_pose(ctx, &c1, points1.data(), points1.shape(0), &c2, points2.data(),
                                                               points2.shape(0), inl.data(), static_cast<int64_t>(inl.size() / 2),
                                                               &g));
          Rigid3d pose;
          std::copy(g.qvec, g.qvec + 4, pose.rotation.wxyz.begin());
          std::copy(g.tvec, g.tvec + 3, pose.translation.begin());
          d["cam2_from_cam1"] = pose;
          return res;
        },
        "points2D1"_a, "points2D2"_a, "camera1"_a, "camera2"_a, "estimation_options"_a = RANSACOptions());
  m.def("cam_from_img",
        [](const py::object& camera, const ArrD& points) {
          CheckPoints(points, "points");
          const b2m_camera c = CameraFromPython(camera);
          ArrD out(std::vector<py::ssize_t>{points.shape(0), 2});
          b2m_ctx* ctx = Engine::Get(0);
          ThrowOnError(ctx, b2m_cam_from_img(ctx, &c, points.data(), points.shape(0), out.mutable_data()));
          return out;
        },
        "camera"_a, "points"_a, "Camera.cam_from_img on an N x 2 point list (R:scene/camera.h), every supported model");
  m.def("squared_sampson_error",
        [](const ArrD& points1, const ArrD& points2, const ArrD& E) {
          CheckSameLength(points1, points2, "two_view_geometry.h:165");
          if (E.size() != 9) throw std::invalid_argument("[bindings.cc] Check Failed: E is 3 x 3");
          py::array_t<double> out(points1.shape(0));
          b2m_ctx* ctx = Engine::Get(0);
          ThrowOnError(ctx, b2m_squared_sampson_error(ctx, points1.data(), points2.data(), points1.shape(0), E.data(),
                                                      out.mutable_data()));
          return out;
        },
        "points2D1"_a, "points2D2"_a, "E"_a);

  // ---- pair generators (exposed for tests and tools) ----
  m.def("exhaustive_pair_blocks",
        [](int n_images, int block_size) {
          py::list out;
          for (const PairList& b : ExhaustivePairBlocks(n_images, block_size)) {
            py::array_t<int32_t> a(std::vector<py::ssize_t>{static_cast<py::ssize_t>(b.size() / 2), 2});
            memcpy(a.mutable_data(), b.data(), b.size() * 4);
            out.append(a);
          }
          return out;
        },
        "n_images"_a, "block_size"_a);
  m.def("comm_image_range",
        [](int n_images, int n_ranks, int rank) {
          int32_t first = 0, count = 0;
          b2m_comm_image_range(n_images, n_ranks, rank, &first, &count);
          return py::make_tuple(first, count);
        },
        "n_images"_a, "n_ranks"_a, "rank"_a, "The contiguous image range rank `rank` uploads (b2m_comm_image_range)");
  m.def("last_pipeline_timing", []() {
    const PipelineTiming t = LastPipelineTiming();
    return py::dict("read_s"_a = t.read_s, "upload_s"_a = t.upload_s, "gpu_s"_a = t.gpu_s, "write_s"_a = t.write_s,
                    "write_wait_s"_a = t.write_wait_s, "total_s"_a = t.total_s, "pairs"_a = t.pairs,
                    "sharded_upload"_a = t.sharded_upload);
  }, "Wall-clock breakdown of the last match_* / verify_matches call of this process");
  m.def("parse_gpu_indices", &ParseGpuIndices, "gpu_index"_a);
  m.def("split_pairs_by_cost",
        [](const ArrI32& pairs, const std::vector<int32_t>& n_feat, int parts) {
          if (pairs.size() % 2) throw std::invalid_argument("[bindings.cc] Check Failed: pairs is N x 2 int32");
          for (py::ssize_t i = 0; i < pairs.size(); ++i)
            if (pairs.data()[i] < 0 || pairs.data()[i] >= static_cast<int32_t>(n_feat.size()))
              throw std::invalid_argument("[bindings.cc] Check Failed: pair index < number of images");
          return SplitPairsByCost(PairList(pairs.data(), pairs.data() + pairs.size()), n_feat, parts);
        },
        "pairs"_a, "n_feat"_a, "parts"_a);
  m.def("spatial_pairs",
        [](const ArrD& prior_t, const std::vector<bool>& has_prior, const SpatialMatchingOptions& options) {
          if (prior_t.ndim() != 2 || prior_t.shape(1) != 3 || prior_t.shape(0) != static_cast<py::ssize_t>(has_prior.size()))
            throw std::invalid_argument("[bindings.cc] Check Failed: prior_t is N x 3 with one flag per row");
          std::vector<std::array<double, 3>> t(has_prior.size());
          for (size_t i = 0; i < t.size(); ++i) t[i] = {prior_t.data()[3 * i], prior_t.data()[3 * i + 1], prior_t.data()[3 * i + 2]};
          const PairList p = SpatialPairs(t, has_prior, options);
          py::array_t<int32_t> a(std::vector<py::ssize_t>{static_cast<py::ssize_t>(p.size() / 2), 2});
          if (!p.empty()) memcpy(a.mutable_data(), p.data(), p.size() * 4);
          return a;
        },
        "prior_t"_a, "has_prior"_a, "options"_a = SpatialMatchingOptions());
  m.def("sequential_pairs",
        [](int n_images, int overlap, bool quadratic_overlap) {
          const PairList p = SequentialPairs(n_images, overlap, quadratic_overlap);
          py::array_t<int32_t> a(std::vector<py::ssize_t>{static_cast<py::ssize_t>(p.size() / 2), 2});
          if (!p.empty()) memcpy(a.mutable_data(), p.data(), p.size() * 4);
          return a;
        },
        "n_images"_a, "overlap"_a, "quadratic_overlap"_a);

  // ---- database ----
  m.def("image_pair_to_pair_id", &ImagePairToPairId, "image_id1"_a, "image_id2"_a);
  m.def("pair_id_to_image_pair", [](int64_t pid) {
    int64_t a, b;
    PairIdToImagePair(pid, &a, &b);
    return py::make_tuple(a, b);
  });
  m.def("sqlite_version", []() { return std::string(sq::api().libversion()); });
  py::class_<Database>(m, "Database")
      .def(py::init<>())
      .def(py::init([](const py::object& path) { return std::make_unique<Database>(FsPath(path)); }), "path"_a)
      .def_static("connect", [](const py::object& path) { return std::make_unique<Database>(FsPath(path)); }, "path"_a)
      .def("open", [](Database& db, const py::object& path) { db.Open(FsPath(path)); }, "path"_a)
      .def("close", &Database::Close)
      .def("__enter__", [](py::object self) { return self; })
      .def("__exit__", [](Database& db, const py::args&) { db.Close(); })
      .def_property_readonly("num_cameras", &Database::NumCameras)
      .def_property_readonly("num_images", &Database::NumImages)
      .def_property_readonly("num_keypoints", &Database::NumKeypoints)
      .def_property_readonly("num_descriptors", &Database::NumDescriptors)
      .def_property_readonly("num_matches", &Database::NumMatches)
      .def_property_readonly("num_inlier_matches", &Database::NumInlierMatches)
      .def_property_readonly("num_matched_image_pairs", &Database::NumMatchedImagePairs)
      .def_property_readonly("num_verified_image_pairs", &Database::NumVerifiedImagePairs)
      .def("num_rows", &Database::NumRows, "table"_a)
      .def("add_camera",
           [](Database& db, int model, int64_t width, int64_t height, const std::vector<double>& params,
              bool prior_focal_length) { return db.AddCamera(model, width, height, params, prior_focal_length); },
           "model"_a, "width"_a, "height"_a, "params"_a, "prior_focal_length"_a = false)
      .def("add_image",
           [](Database& db, const std::string& name, int64_t camera_id, const py::object& prior_t) {
             if (prior_t.is_none()) return db.AddImage(name, camera_id);
             const std::vector<double> t = prior_t.cast<std::vector<double>>();
             if (t.size() != 3) throw std::invalid_argument("[bindings.cc] Check Failed: prior_t has 3 entries");
             return db.AddImage(name, camera_id, std::array<double, 3>{t[0], t[1], t[2]});
           },
           "name"_a, "camera_id"_a, "prior_t"_a = py::none())
      .def("write_keypoints",
           [](Database& db, int64_t image_id, const ArrF32& kp) {
             if (kp.ndim() != 2) throw std::invalid_argument("[bindings.cc] Check Failed: keypoints is a 2-D array");
             db.WriteKeypoints(image_id, kp.data(), kp.shape(0), kp.shape(1));
           },
           "image_id"_a, "keypoints"_a)
      .def("write_descriptors",
           [](Database& db, int64_t image_id, const ArrU8& d) {
             if (d.ndim() != 2) throw std::invalid_argument("[bindings.cc] Check Failed: descriptors is a 2-D array");
             db.WriteDescriptors(image_id, d.data(), d.shape(0), d.shape(1));
           },
           "image_id"_a, "descriptors"_a)
      .def("read_all_images",
           [](Database& db) {
             py::list out;
             for (const ImageRow& r : db.ReadAllImages()) out.append(py::make_tuple(r.image_id, r.name, r.camera_id));
             return out;
           })
      .def("read_camera", [](Database& db, int64_t camera_id) { return CameraToDict(db.ReadCamera(camera_id)); },
           "camera_id"_a)
      // the rest of the reference's Database surface (R:scene/database.h:10-47); cameras are dicts, images are
      // (image_id, name, camera_id) tuples here -- pycolmap.Camera / Image belong to the reconstruction object model
      .def("read_all_cameras",
           [](Database& db) {
             py::list out;
             for (const CameraRow& c : db.ReadAllCameras()) {
               py::dict d = CameraToDict(c);
               d["camera_id"] = c.camera_id;
               out.append(d);
             }
             return out;
           })
      .def("read_image",
           [](Database& db, int64_t image_id) -> py::object {
             ImageRow r;
             if (!db.ReadImage(image_id, &r)) throw std::invalid_argument("[bindings.cc] Check Failed: image exists");
             return py::make_tuple(r.image_id, r.name, r.camera_id);
           },
           "image_id"_a)
      .def("read_image_with_name",
           [](Database& db, const std::string& name) -> py::object {
             ImageRow r;
             if (!db.ReadImageWithName(name, &r)) throw std::invalid_argument("[bindings.cc] Check Failed: image exists");
             return py::make_tuple(r.image_id, r.name, r.camera_id);
           },
           "name"_a)
      .def("num_keypoints_for_image", &Database::NumKeypointsForImage, "image_id"_a)
      .def("num_descriptors_for_image", &Database::NumDescriptorsForImage, "image_id"_a)
      .def("image_pair_to_pair_id", [](Database&, int64_t a, int64_t b) { return ImagePairToPairId(a, b); },
           "image_id1"_a, "image_id2"_a)
      .def("pair_id_to_image_pair",
           [](Database&, int64_t pid) {
             int64_t a, b;
             PairIdToImagePair(pid, &a, &b);
             return py::make_tuple(a, b);
           },
           "pair_id"_a)
      .def("write_camera",
           [](Database& db, const py::object& camera) {
             const b2m_camera c = CameraFromPython(camera);   // validates model id and parameter count
             return db.AddCamera(c.model, c.width, c.height,
                                 std::vector<double>(c.params, c.params + b2m::cam::num_params(c.model)),
                                 c.has_prior_focal_length != 0);
           },
           "camera"_a)
      .def("write_image", [](Database& db, const std::string& name, int64_t camera_id) { return db.AddImage(name, camera_id); },
           "name"_a, "camera_id"_a)
      .def("read_keypoints",
           [](Database& db, int64_t image_id) {
             const KeypointsBlob k = db.ReadKeypoints(image_id);
             py::array_t<float> a(std::vector<py::ssize_t>{static_cast<py::ssize_t>(k.rows), static_cast<py::ssize_t>(k.cols)});
             if (!k.data.empty()) memcpy(a.mutable_data(), k.data.data(), k.data.size() * 4);
             return a;
           },
           "image_id"_a)
      .def("read_descriptors",
           [](Database& db, int64_t image_id) {
             const DescriptorsBlob d = db.ReadDescriptors(image_id);
             py::array_t<uint8_t> a(std::vector<py::ssize_t>{static_cast<py::ssize_t>(d.rows), 128});
             if (!d.data.empty()) memcpy(a.mutable_data(), d.data.data(), d.data.size());
             return a;
           },
           "image_id"_a)
      .def("exists_matches", &Database::ExistsMatches, "image_id1"_a, "image_id2"_a)
      .def("exists_inlier_matches", &Database::ExistsInlierMatches, "image_id1"_a, "image_id2"_a)
      .def("read_matches",
           [](Database& db, int64_t id1, int64_t id2) {
             const std::vector<uint32_t> mm = db.ReadMatches(id1, id2);
             return MatchesToNumpy(mm.data(), static_cast<int64_t>(mm.size() / 2));
           },
           "image_id1"_a, "image_id2"_a)
      .def("read_two_view_geometry",
           [](Database& db, int64_t id1, int64_t id2) -> py::object {
             TwoViewGeometryRow row;
             if (!db.ReadTwoViewGeometry(id1, id2, &row)) return py::none();
             TwoViewGeometry g;
             g.config = static_cast<TwoViewGeometryConfiguration>(row.config);
             g.E = row.E; g.F = row.F; g.H = row.H;
             g.qvec = row.qvec; g.tvec = row.tvec;
             g.inlier_matches = std::move(row.inlier_matches);
             return py::cast(std::move(g));
           },
           "image_id1"_a, "image_id2"_a)
      .def("write_matches",
           [](Database& db, int64_t id1, int64_t id2, const ArrU32& mm) {
             if (mm.size() % 2) throw std::invalid_argument("[bindings.cc] Check Failed: matches is N x 2");
             db.WriteMatches(id1, id2, mm.data(), mm.size() / 2);
           },
           "image_id1"_a, "image_id2"_a, "matches"_a)
      .def("write_two_view_geometry",
           [](Database& db, int64_t id1, int64_t id2, const TwoViewGeometry& g) {
             db.WriteTwoViewGeometry(id1, id2, static_cast<int>(g.config), g.inlier_matches.data(),
                                     static_cast<int64_t>(g.inlier_matches.size() / 2), g.F, g.E, g.H, g.qvec, g.tvec);
           },
           "image_id1"_a, "image_id2"_a, "two_view_geometry"_a)
      .def("clear_matches", &Database::ClearMatches)
      .def("clear_two_view_geometries", &Database::ClearTwoViewGeometries)
      .def("begin", &Database::Begin)
      .def("commit", &Database::Commit)
      .def("rollback", &Database::Rollback);

  py::class_<DatabaseTransaction>(m, "DatabaseTransaction")   // R:scene/database.h:45-46: BEGIN now, COMMIT when released
      .def(py::init<Database*>(), "database"_a, py::keep_alive<1, 2>());

  // ---- low-level context ----
  py::class_<CoreResults>(m, "Results")
      .def("__len__", [](const CoreResults& r) { return r.r ? b2m_results_num_pairs(r.r) : 0; })
      .def_property_readonly("total_matches", [](const CoreResults& r) { return r.r ? b2m_results_total_matches(r.r) : 0; })
      .def_property_readonly("num_verified", [](const CoreResults& r) { return r.r ? b2m_results_num_verified(r.r) : 0; })
      .def("matches", [](const CoreResults& r, int64_t k) {
        const b2m_pair_view v = r.View(k);
        return MatchesToNumpy(v.matches, v.n_matches);
      })
      .def("inlier_matches", [](const CoreResults& r, int64_t k) {
        const b2m_pair_view v = r.View(k);
        return MatchesToNumpy(v.inlier_matches, v.n_inliers);
      })
      .def("two_view_geometry", [](const CoreResults& r, int64_t k) {
        const b2m_pair_view v = r.View(k);
        TwoViewGeometry g;
        g.config = static_cast<TwoViewGeometryConfiguration>(v.config);
        std::copy(v.E, v.E + 9, g.E.begin());
        std::copy(v.F, v.F + 9, g.F.begin());
        std::copy(v.H, v.H + 9, g.H.begin());
        g.inlier_matches.assign(v.inlier_matches, v.inlier_matches + 2 * v.n_inliers);
        std::copy(v.qvec, v.qvec + 4, g.qvec.begin());
        std::copy(v.tvec, v.tvec + 3, g.tvec.begin());
        g.tri_angle = v.tri_angle;
        return g;
      })
      .def("image_pair", [](const CoreResults& r, int64_t k) {
        const b2m_pair_view v = r.View(k);
        return py::make_tuple(v.image1, v.image2);
      })
      .def("free", &CoreResults::Free);

  py::class_<CoreContext>(m, "Context")
      .def(py::init<int, uint64_t, int>(), "device"_a = 0, "seed"_a = 0, "pair_batch"_a = 0)
      .def("close", &CoreContext::Close)
      .def("set_images",
           [](CoreContext& c, const py::list& descs, const py::object& kpts, const py::object& cams) {
             const size_t n = descs.size();
             std::vector<ArrU8> d(n);
             std::vector<ArrF32> k(n);
             std::vector<int32_t> n_feat(n);
             std::vector<const uint8_t*> dptr(n);
             std::vector<const float*> kptr(n);
             for (size_t i = 0; i < n; ++i) {
               d[i] = ArrU8::ensure(descs[i]);
               if (!d[i] || d[i].size() % 128)
                 throw std::invalid_argument("[bindings.cc] Check Failed: descriptors are N x 128 uint8");
               n_feat[i] = static_cast<int32_t>(d[i].size() / 128);
               dptr[i] = d[i].data();
             }
             if (!kpts.is_none()) {
               const py::list kl = kpts;
               if (kl.size() != n) throw std::invalid_argument("[bindings.cc] Check Failed: one keypoint array per image");
               for (size_t i = 0; i < n; ++i) {
                 k[i] = ArrF32::ensure(kl[i]);
                 if (!k[i] || k[i].size() != 2 * static_cast<py::ssize_t>(n_feat[i]))
                   throw std::invalid_argument("[bindings.cc] Check Failed: keypoints.rows == descriptors.rows (N x 2 float32)");
                 kptr[i] = k[i].data();
               }
             }
             std::vector<b2m_camera> cc;
             if (!cams.is_none()) {
               cc = CamerasFromPython(cams);
               if (cc.size() != n) throw std::invalid_argument("[bindings.cc] Check Failed: one camera per image");
             }
             b2m_ctx* ctx = c.Handle();
             int rc;
             {
               py::gil_scoped_release release;
               rc = b2m_set_images(ctx, static_cast<int32_t>(n), n_feat.data(), dptr.data(),
                                   kpts.is_none() ? nullptr : kptr.data(), cams.is_none() ? nullptr : cc.data());
             }
             ThrowOnError(ctx, rc);
           },
           "descriptors"_a, "keypoints"_a = py::none(), "cameras"_a = py::none())
      .def("set_images_device",
           [](CoreContext& c, const ArrI32& n_feat, uint64_t dev_desc_ptr, uint64_t dev_kpts_ptr, const py::object& cams) {
             std::vector<b2m_camera> cc;
             if (!cams.is_none()) cc = CamerasFromPython(cams);
             b2m_ctx* ctx = c.Handle();
             ThrowOnError(ctx, b2m_set_images_device(ctx, static_cast<int32_t>(n_feat.size()), n_feat.data(),
                                                     reinterpret_cast<const void*>(dev_desc_ptr),
                                                     reinterpret_cast<const void*>(dev_kpts_ptr),
                                                     cams.is_none() ? nullptr : cc.data()));
           },
           "n_feat"_a, "dev_desc_ptr"_a, "dev_kpts_ptr"_a = 0, "cameras"_a = py::none())
      // ---- multi-GPU (include/b200match.h: b2m_comm_*, b2m_set_images_sharded) ----
      .def_static("comm_unique_id",
                  []() {
                    b2m_comm_id id;
                    ThrowOnError(nullptr, b2m_comm_get_unique_id(&id));
                    return py::bytes(reinterpret_cast<const char*>(id.bytes), B2M_COMM_ID_BYTES);
                  },
                  "Rank 0: a fresh communicator id (128 bytes) to hand to every rank through the launcher's side channel")
      .def("comm_init_rank",
           [](CoreContext& c, int n_ranks, int rank, const py::bytes& id) {
             const std::string raw = id;
             if (raw.size() != B2M_COMM_ID_BYTES) throw std::invalid_argument("[bindings.cc] Check Failed: id has 128 bytes");
             b2m_comm_id cid;
             memcpy(cid.bytes, raw.data(), B2M_COMM_ID_BYTES);
             b2m_ctx* ctx = c.Handle();
             int rc;
             {
               py::gil_scoped_release release;
               rc = b2m_comm_init_rank(ctx, n_ranks, rank, &cid);
             }
             ThrowOnError(ctx, rc);
           },
           "n_ranks"_a, "rank"_a, "id"_a)
      .def("comm_destroy", [](CoreContext& c) { ThrowOnError(c.Handle(), b2m_comm_destroy(c.Handle())); })
      .def("set_images_sharded",
           [](CoreContext& c, const ArrI32& n_feat, int first_image, int n_local, const py::object& desc_local,
              const py::object& kpts_local, const py::object& cams, bool has_keypoints) {
             // desc_local / kpts_local: numpy arrays (host: [rows x 128] uint8, [rows x 2] float32) or device pointers (int)
             b2m_image_shard sh;
             memset(&sh, 0, sizeof(sh));
             sh.struct_size = sizeof(sh);
             sh.first_image = first_image;
             sh.n_local = n_local;
             sh.has_keypoints = has_keypoints ? 1 : 0;
             ArrU8 d;
             ArrF32 k;
             if (py::isinstance<py::int_>(desc_local)) {
               sh.location = B2M_LOC_DEVICE;
               sh.desc_packed = reinterpret_cast<const void*>(desc_local.cast<uint64_t>());
               if (!kpts_local.is_none()) sh.kpts_packed = reinterpret_cast<const void*>(kpts_local.cast<uint64_t>());
             } else {
               sh.location = B2M_LOC_HOST;
               int64_t rows = 0;
               for (int i = first_image; i < first_image + n_local && i < n_feat.size(); ++i) rows += n_feat.data()[i];
               if (!desc_local.is_none()) {
                 d = ArrU8::ensure(desc_local);
                 if (!d || d.size() != rows * 128)
                   throw std::invalid_argument("[bindings.cc] Check Failed: local descriptors are (sum of local n_feat) x 128 uint8");
                 sh.desc_packed = d.data();
               }
               if (!kpts_local.is_none()) {
                 k = ArrF32::ensure(kpts_local);
                 if (!k || k.size() != rows * 2)
                   throw std::invalid_argument("[bindings.cc] Check Failed: local keypoints are (sum of local n_feat) x 2 float32");
                 sh.kpts_packed = k.data();
               }
             }
             std::vector<b2m_camera> cc;
             if (!cams.is_none()) {
               cc = CamerasFromPython(cams);
               if (static_cast<py::ssize_t>(cc.size()) != n_feat.size())
                 throw std::invalid_argument("[bindings.cc] Check Failed: one camera per image (all images, not only the local ones)");
             }
             b2m_ctx* ctx = c.Handle();
             int rc;
             {
               py::gil_scoped_release release;
               rc = b2m_set_images_sharded(ctx, static_cast<int32_t>(n_feat.size()), n_feat.data(),
                                           cams.is_none() ? nullptr : cc.data(), &sh);
             }
             ThrowOnError(ctx, rc);
           },
           "n_feat"_a, "first_image"_a, "n_local"_a, "descriptors_local"_a, "keypoints_local"_a = py::none(),
           "cameras"_a = py::none(), "has_keypoints"_a = false)
      .def("match_pair",
           [](CoreContext& c, const ArrU8& d1, const ArrU8& d2, const SiftMatchingOptions& options) {
             if (d1.size() % 128 || d2.size() % 128)
               throw std::invalid_argument("[bindings.cc] Check Failed: descriptors are N x 128 uint8");
             const int32_t n1 = static_cast<int32_t>(d1.size() / 128), n2 = static_cast<int32_t>(d2.size() / 128);
             std::vector<uint32_t> out(static_cast<size_t>(std::max(1, n1)) * 2);
             int64_t n = 0;
             const b2m_sift_opts o = ToAbi(options);
             b2m_ctx* ctx = c.Handle();
             int rc;
             {
               py::gil_scoped_release release;
               rc = b2m_match_pair(ctx, d1.data(), n1, d2.data(), n2, &o, out.data(), std::max(1, n1), &n);
             }
             ThrowOnError(ctx, rc);
             return MatchesToNumpy(out.data(), n);
           },
           "descriptors1"_a, "descriptors2"_a, "options"_a = SiftMatchingOptions())
      .def("match_pairs",
           [](CoreContext& c, const ArrI32& pairs, const SiftMatchingOptions& sift, const py::object& tvg) {
             if (pairs.size() % 2) throw std::invalid_argument("[bindings.cc] Check Failed: pairs is N x 2 int32");
             const b2m_sift_opts so = ToAbi(sift);
             b2m_tvg_opts to;
             const bool verify = !tvg.is_none();
             if (verify) to = ToAbi(tvg.cast<TwoViewGeometryOptions>());
             auto res = std::make_unique<CoreResults>();
             b2m_ctx* ctx = c.Handle();
             int rc = B2M_OK;
             RunInterruptible([&] {
               rc = b2m_match_pairs(ctx, pairs.data(), pairs.size() / 2, &so, verify ? &to : nullptr, &res->r);
             });
             ThrowOnError(ctx, rc);
             return res;
           },
           "pairs"_a, "sift_options"_a = SiftMatchingOptions(), "verification_options"_a = py::none())
      .def("estimate_two_view_geometry",
           [](CoreContext& c, const py::object& camera1, const ArrD& points1, const py::object& camera2,
              const ArrD& points2, const py::object& matches, const TwoViewGeometryOptions& options) {
             return EstimateTvg([&c] { return c.Handle(); }, camera1, points1, camera2, points2, matches, options);
           },
           "camera1"_a, "points1"_a, "camera2"_a, "points2"_a, "matches"_a = py::none(),
           "options"_a = TwoViewGeometryOptions())
      .def("debug_five_point",
           [](CoreContext& c, const ArrD& nullspaces) {
             if (nullspaces.size() % 36) throw std::invalid_argument("[bindings.cc] Check Failed: null spaces are n x 4 x 9");
             const int64_t n = nullspaces.size() / 36;
             ArrD models(std::vector<py::ssize_t>{static_cast<py::ssize_t>(n), 10, 9});
             py::array_t<int32_t> counts(static_cast<py::ssize_t>(n));
             ThrowOnError(c.Handle(), b2m_debug_five_point(c.Handle(), nullspaces.data(), n, models.mutable_data(),
                                                           counts.mutable_data()));
             return py::make_tuple(models, counts);
           },
           "nullspaces"_a, "The device 5-point solver (one warp per 4-D null space): (models [n, 10, 9], counts [n])")
      .def("stats", [](CoreContext& c) {
        b2m_stats s;
        memset(&s, 0, sizeof(s));
        s.struct_size = sizeof(s);
        ThrowOnError(c.Handle(), b2m_get_stats(c.Handle(), &s));
        return py::dict("kernel_launches"_a = s.kernel_launches, "match_tiles"_a = s.match_tiles,
                        "last_match_ms"_a = s.last_match_ms, "last_verify_ms"_a = s.last_verify_ms,
                        "last_total_ms"_a = s.last_total_ms, "last_k1_ms"_a = s.last_k1_ms,
                        "last_k1_launches"_a = s.last_k1_launches,
                        "k1_dir1_mode"_a = s.k1_dir1_mode, "last_allgather_ms"_a = s.last_allgather_ms,
                        "last_allgather_bytes"_a = s.last_allgather_bytes, "last_upload_ms"_a = s.last_upload_ms,
                        "verify_models_scored"_a = py::make_tuple(s.verify_models_scored[0], s.verify_models_scored[1],
                                                                  s.verify_models_scored[2]),
                        "verify_residuals"_a = py::make_tuple(s.verify_residuals[0], s.verify_residuals[1],
                                                              s.verify_residuals[2]),
                        "comm_size"_a = s.comm_size, "comm_rank"_a = s.comm_rank);
      })
      .def("reset_stats", [](CoreContext& c) { ThrowOnError(c.Handle(), b2m_reset_stats(c.Handle()));
      });

  // test hook for the PyWait logic: blocks `seconds` on the worker thread like a long GPU call would
  m.def("_sleep_interruptible", [](double seconds) {
    RunInterruptible([seconds] { std::this_thread::sleep_for(std::chrono::duration<double>(seconds)); });
  });

  // destroy the per-device contexts of the pipeline / estimator functions at interpreter exit
  m.add_object("_cleanup", py::capsule([]() { Engine::DestroyAll(); }));
}
