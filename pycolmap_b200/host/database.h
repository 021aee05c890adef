// database.h -- COLMAP SQLite database access for the matching / verification controllers
// (SURVEY.md row S1).  Schema and value conventions: U:scene/database.cc (COLMAP 3.9.1):
//   pair_id = id1 * 2147483647 + id2 with id1 < id2; matches / two-view geometries of a pair given as
//   (id1 > id2) are stored swapped (columns exchanged, F and E transposed, H inverted).
// The reference binds open/close, the counters, read_two_view_geometry and image_pair_to_pair_id
// (R:scene/database.h:10-47); the readers / writers the controllers need are unbound upstream members
// and are implemented here.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "sqlite_dyn.h"

namespace b2mh {

constexpr int64_t kMaxNumImages = 2147483647;  // U:scene/database.h kMaxNumImages

int64_t ImagePairToPairId(int64_t image_id1, int64_t image_id2);  // R:scene/database.h:28-29
void PairIdToImagePair(int64_t pair_id, int64_t* image_id1, int64_t* image_id2);

struct ImageRow {
  int64_t image_id = 0;
  std::string name;
  int64_t camera_id = 0;
};

struct CameraRow {
  int64_t camera_id = 0;
  int model = 0;  // COLMAP camera model id (0 SIMPLE_PINHOLE, 1 PINHOLE, 2 SIMPLE_RADIAL, ...)
  int64_t width = 0, height = 0;
  std::vector<double> params;
  bool has_prior_focal_length = false;
};

struct KeypointsBlob {
  int64_t rows = 0, cols = 0;  // cols in {2, 4, 6}: x, y [, scale, orientation | a11, a12, a21, a22]
  std::vector<float> data;     // row-major
};

struct DescriptorsBlob {
  int64_t rows = 0, cols = 128;
  std::vector<uint8_t> data;  // row-major
};

using Mat3 = std::array<double, 9>;  // row-major

struct TwoViewGeometryRow {
  int config = 0;
  Mat3 F{}, E{}, H{};
  std::vector<uint32_t> inlier_matches;  // [n x 2]
  std::array<double, 4> qvec{1.0, 0.0, 0.0, 0.0};  // cam2_from_cam1 rotation (w, x, y, z)
  std::array<double, 3> tvec{0.0, 0.0, 0.0};
};

// Pose algebra on (qvec = (w, x, y, z), tvec): rotation matrix, inverse pose (TwoViewGeometry::Invert).
Mat3 QuatToRotation(const std::array<double, 4>& q);
void InvertPose(std::array<double, 4>* qvec, std::array<double, 3>* tvec);

bool Invert3x3(const Mat3& m, Mat3* out);  // false when singular (|det| < 1e-300)
Mat3 Transposed(const Mat3& m);

class Database {
 public:
  Database() = default;
  explicit Database(const std::string& path) { Open(path); }
  ~Database() { Close(); }
  Database(const Database&) = delete;
  Database& operator=(const Database&) = delete;

  void Open(const std::string& path);  // creates the COLMAP tables when missing
  void Close();
  bool IsOpen() const { return db_ != nullptr; }

  // counters (R:scene/database.h:18-27)
  int64_t NumCameras() { return Scalar("SELECT COUNT(*) FROM cameras"); }
  int64_t NumImages() { return Scalar("SELECT COUNT(*) FROM images"); }
  int64_t NumKeypoints() { return Scalar("SELECT COALESCE(SUM(rows),0) FROM keypoints"); }
  int64_t NumDescriptors() { return Scalar("SELECT COALESCE(SUM(rows),0) FROM descriptors"); }
  int64_t NumMatches() { return Scalar("SELECT COALESCE(SUM(rows),0) FROM matches"); }
  int64_t NumInlierMatches() { return Scalar("SELECT COALESCE(SUM(rows),0) FROM two_view_geometries"); }
  int64_t NumMatchedImagePairs() { return Scalar("SELECT COUNT(*) FROM matches WHERE rows > 0"); }
  int64_t NumVerifiedImagePairs() { return Scalar("SELECT COUNT(*) FROM two_view_geometries WHERE rows > 0"); }
  int64_t NumRows(const std::string& table);  // raw row count of `matches` / `two_view_geometries`

  // feature side (written by the extractor upstream; writers are here to build test / bench databases)
  int64_t AddCamera(int model, int64_t width, int64_t height, const std::vector<double>& params,
                    bool prior_focal_length);
  int64_t AddImage(const std::string& name, int64_t camera_id);
  // with a location prior (prior_tx, prior_ty, prior_tz): GPS (latitude, longitude, altitude) or Cartesian
  int64_t AddImage(const std::string& name, int64_t camera_id, const std::array<double, 3>& prior_t);
  // location prior of every image, in ReadAllImages order; has_prior[i] == false where the columns are NULL
  void ReadLocationPriors(std::vector<std::array<double, 3>>* prior_t, std::vector<bool>* has_prior);
  void WriteKeypoints(int64_t image_id, const float* data, int64_t rows, int64_t cols);
  void WriteDescriptors(int64_t image_id, const uint8_t* data, int64_t rows, int64_t cols);

  std::vector<ImageRow> ReadAllImages();  // ordered by image_id
  bool ReadImage(int64_t image_id, ImageRow* out);
  bool ReadImageWithName(const std::string& name, ImageRow* out);
  std::vector<CameraRow> ReadAllCameras();
  int64_t NumKeypointsForImage(int64_t image_id);
  int64_t NumDescriptorsForImage(int64_t image_id);
  CameraRow ReadCamera(int64_t camera_id);
  KeypointsBlob ReadKeypoints(int64_t image_id);
  DescriptorsBlob ReadDescriptors(int64_t image_id);

  // result side (U:scene/database.cc WriteMatches / WriteTwoViewGeometry / Exists* / Read*)
  bool ExistsMatches(int64_t id1, int64_t id2);
  bool ExistsInlierMatches(int64_t id1, int64_t id2);
  std::unordered_set<int64_t> ExistingPairIds(const std::string& table);
  std::vector<uint32_t> ReadMatches(int64_t id1, int64_t id2);  // [n x 2] in (id1, id2) orientation
  bool ReadTwoViewGeometry(int64_t id1, int64_t id2, TwoViewGeometryRow* out);
  void WriteMatches(int64_t id1, int64_t id2, const uint32_t* matches, int64_t n);
  void WriteTwoViewGeometry(int64_t id1, int64_t id2, int config, const uint32_t* inlier_matches, int64_t n,
                            const Mat3& F, const Mat3& E, const Mat3& H,
                            const std::array<double, 4>& qvec = {1.0, 0.0, 0.0, 0.0},
                            const std::array<double, 3>& tvec = {0.0, 0.0, 0.0});
  void ClearTwoViewGeometries() { Exec("DELETE FROM two_view_geometries"); }
  void ClearMatches() { Exec("DELETE FROM matches"); }

  void Begin() { Exec("BEGIN"); }
  void Commit() { Exec("COMMIT"); }
  void Rollback() { Exec("ROLLBACK"); }

 private:
  struct Stmt;  // RAII prepared statement
  void Exec(const char* sql);
  int64_t Scalar(const char* sql);
  [[noreturn]] void Fail(const char* what);
  sq::sqlite3* db_ = nullptr;
  // prepared statements of the per-pair calls (exists / read / write of matches and two-view geometries), keyed by the
  // address of their SQL literal and reused across calls: an exhaustive run issues hundreds of thousands of them
  std::unordered_map<const void*, sq::sqlite3_stmt*> stmt_cache_;
};

// DatabaseTransaction: BEGIN in the constructor, COMMIT in the destructor (ROLLBACK when unwinding).
class DatabaseTransaction {
 public:
  explicit DatabaseTransaction(Database* db);
  ~DatabaseTransaction();

 private:
  Database* db_;
  int exceptions_;
};

}  // namespace b2mh
