#include "database.h"

#include <cmath>
#include <cstring>
#include <exception>
#include <stdexcept>

namespace b2mh {
namespace {

const char* kCreateSql =
    "CREATE TABLE IF NOT EXISTS cameras (camera_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, "
    "model INTEGER NOT NULL, width INTEGER NOT NULL, height INTEGER NOT NULL, params BLOB, "
    "prior_focal_length INTEGER NOT NULL);"
    "CREATE TABLE IF NOT EXISTS images (image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, "
    "name TEXT NOT NULL UNIQUE, camera_id INTEGER NOT NULL, prior_qw REAL, prior_qx REAL, prior_qy REAL, "
    "prior_qz REAL, prior_tx REAL, prior_ty REAL, prior_tz REAL, "
    "CONSTRAINT image_id_check CHECK(image_id >= 0 and image_id < 2147483647), "
    "FOREIGN KEY(camera_id) REFERENCES cameras(camera_id));"
    "CREATE TABLE IF NOT EXISTS keypoints (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, "
    "cols INTEGER NOT NULL, data BLOB, FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);"
    "CREATE TABLE IF NOT EXISTS descriptors (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, "
    "cols INTEGER NOT NULL, data BLOB, FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);"
    "CREATE TABLE IF NOT EXISTS matches (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, "
    "cols INTEGER NOT NULL, data BLOB);"
    "CREATE TABLE IF NOT EXISTS two_view_geometries (pair_id INTEGER PRIMARY KEY NOT NULL, "
    "rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB, config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB, "
    "qvec BLOB, tvec BLOB);"
    "CREATE UNIQUE INDEX IF NOT EXISTS index_name ON images(name);";

bool AllZero(const Mat3& m) {
  for (double v : m)
    if (v != 0.0) return false;
  return true;
}

}  // namespace

int64_t ImagePairToPairId(int64_t a, int64_t b) {
  if (a > b) std::swap(a, b);
  return a * kMaxNumImages + b;
}

void PairIdToImagePair(int64_t pair_id, int64_t* a, int64_t* b) {
  *a = pair_id / kMaxNumImages;
  *b = pair_id % kMaxNumImages;
}

bool Invert3x3(const Mat3& m, Mat3* out) {
  const double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c0 + m[1] * c1 + m[2] * c2;
  if (!(std::fabs(det) > 1e-300)) return false;
  const double r = 1.0 / det;
  (*out)[0] = c0 * r;
  (*out)[1] = (m[2] * m[7] - m[1] * m[8]) * r;
  (*out)[2] = (m[1] * m[5] - m[2] * m[4]) * r;
  (*out)[3] = c1 * r;
  (*out)[4] = (m[0] * m[8] - m[2] * m[6]) * r;
  (*out)[5] = (m[2] * m[3] - m[0] * m[5]) * r;
  (*out)[6] = c2 * r;
  (*out)[7] = (m[1] * m[6] - m[0] * m[7]) * r;
  (*out)[8] = (m[0] * m[4] - m[1] * m[3]) * r;
  return true;
}

Mat3 QuatToRotation(const std::array<double, 4>& q) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  return Mat3{1 - 2 * (y * y + z * z), 2 * (x * y - z * w),     2 * (x * z + y * w),
              2 * (x * y + z * w),     1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
              2 * (x * z - y * w),     2 * (y * z + x * w),     1 - 2 * (x * x + y * y)};
}

void InvertPose(std::array<double, 4>* q, std::array<double, 3>* t) {
  // the identity (no pose estimated) stays bit-for-bit the identity: no -0.0 in the database
  if ((*q)[0] == 1.0 && (*q)[1] == 0.0 && (*q)[2] == 0.0 && (*q)[3] == 0.0 && (*t)[0] == 0.0 && (*t)[1] == 0.0 && (*t)[2] == 0.0)
    return;
  const Mat3 R = QuatToRotation(*q);
  const std::array<double, 3> o = *t;
  for (int i = 0; i < 3; ++i) (*t)[i] = -(R[i] * o[0] + R[3 + i] * o[1] + R[6 + i] * o[2]);  // -R^T t
  (*q)[1] = -(*q)[1];
  (*q)[2] = -(*q)[2];
  (*q)[3] = -(*q)[3];
}

Mat3 Transposed(const Mat3& m) { return Mat3{m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]}; }

// ---- RAII prepared statement ---------------------------------------------------------------------
struct Database::Stmt {
  // cached: `sql` is a string literal (its address is the key); the statement is reset, not finalized, when this
  // handle goes out of scope, and lives until Close().  One handle per SQL text at a time (leaf functions only).
  Stmt(Database* d, const char* sql, bool cached = false) : db(d), cached_(cached) {
    if (!d->db_) throw std::runtime_error("[database.cc] Check Failed: database is open");
    if (cached) {
      auto it = d->stmt_cache_.find(sql);
      if (it != d->stmt_cache_.end()) {
        st = it->second;
        return;
      }
    }
    if (sq::api().prepare_v2(d->db_, sql, -1, &st, nullptr) != sq::kOk) d->Fail(sql);
    if (cached) d->stmt_cache_[sql] = st;
  }
  ~Stmt() {
    if (!st) return;
    if (cached_) {
      sq::api().reset(st);
      sq::api().clear_bindings(st);
    } else {
      sq::api().finalize(st);
    }
  }
  Stmt(const Stmt&) = delete;
  Stmt& operator=(const Stmt&) = delete;
  void I64(int i, int64_t v) { Check(sq::api().bind_int64(st, i, v)); }
  void F64(int i, double v) { Check(sq::api().bind_double(st, i, v)); }
  void Null(int i) { Check(sq::api().bind_null(st, i)); }
  void Text(int i, const std::string& s) {
    Check(sq::api().bind_text(st, i, s.c_str(), static_cast<int>(s.size()), sq::Transient()));
  }
  // the buffer must stay alive until Step() returns (no copy: SQLITE_STATIC)
  void Blob(int i, const void* p, size_t n) {
    static const char kEmpty = 0;
    Check(sq::api().bind_blob64(st, i, n ? p : &kEmpty, n, nullptr));
  }
  bool Step() {
    const int rc = sq::api().step(st);
    if (rc == sq::kRow) return true;
    if (rc != sq::kDone) db->Fail("sqlite3_step");
    return false;
  }
  int64_t ColI64(int i) { return sq::api().column_int64(st, i); }
  double ColF64(int i) { return sq::api().column_double(st, i); }
  bool ColIsNull(int i) { return sq::api().column_type(st, i) == sq::kTypeNull; }
  std::string ColText(int i) {
    const unsigned char* t = sq::api().column_text(st, i);
    return t ? std::string(reinterpret_cast<const char*>(t)) : std::string();
  }
  template <typename T>
  std::vector<T> ColBlob(int i) {
    const void* p = sq::api().column_blob(st, i);
    const size_t n = static_cast<size_t>(sq::api().column_bytes(st, i));
    std::vector<T> v(n / sizeof(T));
    if (p && !v.empty()) memcpy(v.data(), p, v.size() * sizeof(T));
    return v;
  }
  void Check(int rc) {
    if (rc != sq::kOk) db->Fail("sqlite3_bind");
  }
  Database* db;
  sq::sqlite3_stmt* st = nullptr;
  bool cached_ = false;
};

void Database::Fail(const char* what) {
  std::string msg = std::string("[database.cc] SQLite error (") + what + "): ";
  msg += db_ ? sq::api().errmsg(db_) : "database is closed";
  throw std::runtime_error(msg);
}

void Database::Open(const std::string& path) {
  Close();
  if (sq::api().open_v2(path.c_str(), &db_, sq::kOpenReadWrite | sq::kOpenCreate, nullptr) != sq::kOk) {
    std::string msg = "[database.cc] cannot open database " + path + ": " + (db_ ? sq::api().errmsg(db_) : "?");
    if (db_) sq::api().close(db_);
    db_ = nullptr;
    throw std::runtime_error(msg);
  }
  // as upstream's Database::Open (U:scene/database.cc): do not wait for the operating system to flush, keep the
  // rollback journal and temporary tables in memory.  A matching run rewrites hundreds of MB of blobs; the default
  // (synchronous=FULL, journal file on disk) costs 2-3x on a real file system.
  Exec("PRAGMA synchronous=OFF");
  Exec("PRAGMA journal_mode=MEMORY");
  Exec("PRAGMA temp_store=MEMORY");
  Exec(kCreateSql);
}

void Database::Close() {
  for (auto& kv : stmt_cache_) sq::api().finalize(kv.second);
  stmt_cache_.clear();
  if (db_) sq::api().close(db_);
  db_ = nullptr;
}

void Database::Exec(const char* sql) {
  if (!db_) throw std::runtime_error("[database.cc] Check Failed: database is open");
  char* err = nullptr;
  if (sq::api().exec(db_, sql, nullptr, nullptr, &err) != sq::kOk) {
    std::string msg = std::string("[database.cc] SQLite error: ") + (err ? err : "?");
    if (err) sq::api().free(err);
    throw std::runtime_error(msg);
  }
}

int64_t Database::Scalar(const char* sql) {
  Stmt s(this, sql);
  return s.Step() ? s.ColI64(0) : 0;
}

int64_t Database::NumRows(const std::string& table) {
  if (table != "matches" && table != "two_view_geometries")
    throw std::invalid_argument("[database.cc] Check Failed: table is matches or two_view_geometries");
  return Scalar(("SELECT COUNT(*) FROM " + table).c_str());
}

int64_t Database::AddCamera(int model, int64_t width, int64_t height, const std::vector<double>& params,
                            bool prior_focal_length) {
  Stmt s(this, "INSERT INTO cameras VALUES (NULL, ?, ?, ?, ?, ?)");
  s.I64(1, model); s.I64(2, width); s.I64(3, height);
  s.Blob(4, params.data(), params.size() * sizeof(double));
  s.I64(5, prior_focal_length ? 1 : 0);
  s.Step();
  return sq::api().last_insert_rowid(db_);
}

int64_t Database::AddImage(const std::string& name, int64_t camera_id) {
  Stmt s(this, "INSERT INTO images VALUES (NULL, ?, ?, NULL, NULL, NULL, NULL, NULL, NULL, NULL)");
  s.Text(1, name); s.I64(2, camera_id);
  s.Step();
  return sq::api().last_insert_rowid(db_);
}

int64_t Database::AddImage(const std::string& name, int64_t camera_id, const std::array<double, 3>& t) {
  Stmt s(this, "INSERT INTO images VALUES (NULL, ?, ?, NULL, NULL, NULL, NULL, ?, ?, ?)");
  s.Text(1, name); s.I64(2, camera_id);
  for (int k = 0; k < 3; ++k) s.F64(3 + k, t[k]);
  s.Step();
  return sq::api().last_insert_rowid(db_);
}

void Database::ReadLocationPriors(std::vector<std::array<double, 3>>* prior_t, std::vector<bool>* has_prior) {
  prior_t->clear();
  has_prior->clear();
  Stmt s(this, "SELECT prior_tx, prior_ty, prior_tz FROM images ORDER BY image_id");
  while (s.Step()) {
    const bool ok = !s.ColIsNull(0) && !s.ColIsNull(1) && !s.ColIsNull(2);
    prior_t->push_back({ok ? s.ColF64(0) : 0.0, ok ? s.ColF64(1) : 0.0, ok ? s.ColF64(2) : 0.0});
    has_prior->push_back(ok);
  }
}

void Database::WriteKeypoints(int64_t image_id, const float* data, int64_t rows, int64_t cols) {
  if (cols != 2 && cols != 4 && cols != 6)
    throw std::invalid_argument("[database.cc] Check Failed: keypoints have 2, 4 or 6 columns");
  Stmt s(this, "INSERT OR REPLACE INTO keypoints VALUES (?, ?, ?, ?)");
  s.I64(1, image_id); s.I64(2, rows); s.I64(3, cols);
  s.Blob(4, data, static_cast<size_t>(rows * cols) * sizeof(float));
  s.Step();
}

void Database::WriteDescriptors(int64_t image_id, const uint8_t* data, int64_t rows, int64_t cols) {
  if (cols != 128) throw std::invalid_argument("[database.cc] Check Failed: descriptors.cols() == 128");
  Stmt s(this, "INSERT OR REPLACE INTO descriptors VALUES (?, ?, ?, ?)");
  s.I64(1, image_id); s.I64(2, rows); s.I64(3, cols);
  s.Blob(4, data, static_cast<size_t>(rows * cols));
  s.Step();
}

std::vector<ImageRow> Database::ReadAllImages() {
  std::vector<ImageRow> out;
  Stmt s(this, "SELECT image_id, name, camera_id FROM images ORDER BY image_id");
  while (s.Step()) out.push_back(ImageRow{s.ColI64(0), s.ColText(1), s.ColI64(2)});
  return out;
}

bool Database::ReadImage(int64_t image_id, ImageRow* out) {
  Stmt s(this, "SELECT image_id, name, camera_id FROM images WHERE image_id = ?");
  s.I64(1, image_id);
  if (!s.Step()) return false;
  *out = ImageRow{s.ColI64(0), s.ColText(1), s.ColI64(2)};
  return true;
}

bool Database::ReadImageWithName(const std::string& name, ImageRow* out) {
  Stmt s(this, "SELECT image_id, name, camera_id FROM images WHERE name = ?");
  s.Text(1, name);
  if (!s.Step()) return false;
  *out = ImageRow{s.ColI64(0), s.ColText(1), s.ColI64(2)};
  return true;
}

std::vector<CameraRow> Database::ReadAllCameras() {
  std::vector<int64_t> ids;
  {
    Stmt s(this, "SELECT camera_id FROM cameras ORDER BY camera_id");
    while (s.Step()) ids.push_back(s.ColI64(0));
  }
  std::vector<CameraRow> out;
  for (int64_t id : ids) out.push_back(ReadCamera(id));
  return out;
}

int64_t Database::NumKeypointsForImage(int64_t image_id) {
  Stmt s(this, "SELECT rows FROM keypoints WHERE image_id = ?");
  s.I64(1, image_id);
  return s.Step() ? s.ColI64(0) : 0;
}

int64_t Database::NumDescriptorsForImage(int64_t image_id) {
  Stmt s(this, "SELECT rows FROM descriptors WHERE image_id = ?");
  s.I64(1, image_id);
  return s.Step() ? s.ColI64(0) : 0;
}

CameraRow Database::ReadCamera(int64_t camera_id) {
  Stmt s(this, "SELECT model, width, height, params, prior_focal_length FROM cameras WHERE camera_id = ?");
  s.I64(1, camera_id);
  if (!s.Step())
    throw std::invalid_argument("[database.cc] Check Failed: camera " + std::to_string(camera_id) + " exists");
  CameraRow c;
  c.camera_id = camera_id;
  c.model = static_cast<int>(s.ColI64(0));
  c.width = s.ColI64(1);
  c.height = s.ColI64(2);
  c.params = s.ColBlob<double>(3);
  c.has_prior_focal_length = s.ColI64(4) != 0;
  return c;
}

KeypointsBlob Database::ReadKeypoints(int64_t image_id) {
  KeypointsBlob k;
  k.cols = 2;
  Stmt s(this, "SELECT rows, cols, data FROM keypoints WHERE image_id = ?");
  s.I64(1, image_id);
  if (!s.Step() || s.ColI64(0) == 0) return k;
  k.rows = s.ColI64(0);
  k.cols = s.ColI64(1);
  k.data = s.ColBlob<float>(2);
  if (static_cast<int64_t>(k.data.size()) != k.rows * k.cols)
    throw std::runtime_error("[database.cc] Check Failed: keypoints blob size == rows * cols * 4");
  return k;
}

DescriptorsBlob Database::ReadDescriptors(int64_t image_id) {
  DescriptorsBlob d;
  Stmt s(this, "SELECT rows, cols, data FROM descriptors WHERE image_id = ?");
  s.I64(1, image_id);
  if (!s.Step() || s.ColI64(0) == 0) return d;
  d.rows = s.ColI64(0);
  d.cols = s.ColI64(1);
  d.data = s.ColBlob<uint8_t>(2);
  if (d.cols != 128 || static_cast<int64_t>(d.data.size()) != d.rows * d.cols)
    throw std::runtime_error("[database.cc] Check Failed: descriptors are rows x 128 uint8");
  return d;
}

bool Database::ExistsMatches(int64_t id1, int64_t id2) {
  Stmt s(this, "SELECT 1 FROM matches WHERE pair_id = ?", /*cached=*/true);
  s.I64(1, ImagePairToPairId(id1, id2));
  return s.Step();
}

bool Database::ExistsInlierMatches(int64_t id1, int64_t id2) {
  Stmt s(this, "SELECT 1 FROM two_view_geometries WHERE pair_id = ?", /*cached=*/true);
  s.I64(1, ImagePairToPairId(id1, id2));
  return s.Step();
}

std::unordered_set<int64_t> Database::ExistingPairIds(const std::string& table) {
  if (table != "matches" && table != "two_view_geometries")
    throw std::invalid_argument("[database.cc] Check Failed: table is matches or two_view_geometries");
  std::unordered_set<int64_t> out;
  Stmt s(this, ("SELECT pair_id FROM " + table).c_str());
  while (s.Step()) out.insert(s.ColI64(0));
  return out;
}

namespace {
void SwapColumns(std::vector<uint32_t>* m) {
  for (size_t i = 0; i + 1 < m->size(); i += 2) std::swap((*m)[i], (*m)[i + 1]);
}
}  // namespace

std::vector<uint32_t> Database::ReadMatches(int64_t id1, int64_t id2) {
  Stmt s(this, "SELECT rows, cols, data FROM matches WHERE pair_id = ?", /*cached=*/true);
  s.I64(1, ImagePairToPairId(id1, id2));
  if (!s.Step() || s.ColI64(0) == 0) return {};
  std::vector<uint32_t> m = s.ColBlob<uint32_t>(2);
  m.resize(static_cast<size_t>(s.ColI64(0)) * 2);
  if (id1 > id2) SwapColumns(&m);
  return m;
}

bool Database::ReadTwoViewGeometry(int64_t id1, int64_t id2, TwoViewGeometryRow* out) {
  Stmt s(this, "SELECT rows, cols, data, config, F, E, H, qvec, tvec FROM two_view_geometries WHERE pair_id = ?", /*cached=*/true);
  s.I64(1, ImagePairToPairId(id1, id2));
  if (!s.Step()) return false;
  TwoViewGeometryRow g;
  const int64_t rows = s.ColI64(0);
  if (rows > 0) {
    g.inlier_matches = s.ColBlob<uint32_t>(2);
    g.inlier_matches.resize(static_cast<size_t>(rows) * 2);
  }
  g.config = static_cast<int>(s.ColI64(3));
  auto mat = [&](int col, Mat3* m) {
    const std::vector<double> v = s.ColBlob<double>(col);
    if (v.size() == 9) std::copy(v.begin(), v.end(), m->begin());
  };
  mat(4, &g.F); mat(5, &g.E); mat(6, &g.H);
  {
    const std::vector<double> q = s.ColBlob<double>(7), t = s.ColBlob<double>(8);
    if (q.size() == 4) std::copy(q.begin(), q.end(), g.qvec.begin());
    if (t.size() == 3) std::copy(t.begin(), t.end(), g.tvec.begin());
  }
  if (id1 > id2) {  // TwoViewGeometry::Invert (R:estimators/two_view_geometry.h:92)
    InvertPose(&g.qvec, &g.tvec);
    SwapColumns(&g.inlier_matches);
    g.F = Transposed(g.F);
    g.E = Transposed(g.E);
    Mat3 inv;
    if (!AllZero(g.H) && Invert3x3(g.H, &inv)) g.H = inv;
  }
  *out = std::move(g);
  return true;
}

void Database::WriteMatches(int64_t id1, int64_t id2, const uint32_t* matches, int64_t n) {
  std::vector<uint32_t> swapped;
  if (id1 > id2 && n > 0) {
    swapped.assign(matches, matches + 2 * n);
    SwapColumns(&swapped);
    matches = swapped.data();
  }
  Stmt s(this, "INSERT OR REPLACE INTO matches VALUES (?, ?, 2, ?)", /*cached=*/true);
  s.I64(1, ImagePairToPairId(id1, id2)); s.I64(2, n);
  s.Blob(3, matches, static_cast<size_t>(n) * 8);
  s.Step();
}

void Database::WriteTwoViewGeometry(int64_t id1, int64_t id2, int config, const uint32_t* inlier_matches, int64_t n,
                                    const Mat3& F_in, const Mat3& E_in, const Mat3& H_in,
                                    const std::array<double, 4>& qvec_in, const std::array<double, 3>& tvec_in) {
  std::vector<uint32_t> swapped;
  Mat3 F = F_in, E = E_in, H = H_in;
  std::array<double, 4> q = qvec_in;
  std::array<double, 3> t = tvec_in;
  if (id1 > id2) {  // store in the id1 < id2 frame
    InvertPose(&q, &t);
    if (n > 0) {
      swapped.assign(inlier_matches, inlier_matches + 2 * n);
      SwapColumns(&swapped);
      inlier_matches = swapped.data();
    }
    F = Transposed(F);
    E = Transposed(E);
    Mat3 inv;
    if (!AllZero(H) && Invert3x3(H, &inv)) H = inv;
  }
  const double* qvec = q.data();
  const double* tvec = t.data();
  Stmt s(this, "INSERT OR REPLACE INTO two_view_geometries VALUES (?, ?, 2, ?, ?, ?, ?, ?, ?, ?)", /*cached=*/true);
  s.I64(1, ImagePairToPairId(id1, id2)); s.I64(2, n);
  s.Blob(3, inlier_matches, static_cast<size_t>(n) * 8);
  s.I64(4, config);
  s.Blob(5, F.data(), 72); s.Blob(6, E.data(), 72); s.Blob(7, H.data(), 72);
  s.Blob(8, qvec, 32); s.Blob(9, tvec, 24);
  s.Step();
}

DatabaseTransaction::DatabaseTransaction(Database* db) : db_(db), exceptions_(std::uncaught_exceptions()) {
  db_->Begin();
}

DatabaseTransaction::~DatabaseTransaction() {
  try {
    if (std::uncaught_exceptions() > exceptions_) db_->Rollback(); else db_->Commit();
  } catch (...) {
  }
}

}  // namespace b2mh
