"""COLMAP SQLite database access for the matching / verification pipelines.

Schema: U:scene/database.cc (COLMAP 3.9.1).  The reference binds only a few members
(R:scene/database.h:10-47: open/close, counts, read_two_view_geometry, image_pair_to_pair_id);
the descriptor / keypoint / match readers and writers the pipelines need are implemented here with
the standard-library `sqlite3` (no sqlite3.h in this image).
"""
import os
import sqlite3

import numpy as np

MAX_NUM_IMAGES = 2147483647  # kMaxNumImages: pair_id = id1 * kMaxNumImages + id2, id1 < id2

# COLMAP camera model id -> number of parameters, for the models the verifier supports (U:sensor/models.h):
# all eleven models of COLMAP 3.9.1.
CAMERA_MODEL_NUM_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8, 5: 8, 6: 12, 7: 5, 8: 4, 9: 5, 10: 12}
CAMERA_MODEL_IDS = {"SIMPLE_PINHOLE": 0, "PINHOLE": 1, "SIMPLE_RADIAL": 2, "RADIAL": 3, "OPENCV": 4,
                    "OPENCV_FISHEYE": 5, "FULL_OPENCV": 6, "FOV": 7, "SIMPLE_RADIAL_FISHEYE": 8,
                    "RADIAL_FISHEYE": 9, "THIN_PRISM_FISHEYE": 10}

CREATE_SQL = """
CREATE TABLE IF NOT EXISTS cameras (camera_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, model INTEGER NOT NULL,
    width INTEGER NOT NULL, height INTEGER NOT NULL, params BLOB, prior_focal_length INTEGER NOT NULL);
CREATE TABLE IF NOT EXISTS images (image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE,
    camera_id INTEGER NOT NULL, prior_qw REAL, prior_qx REAL, prior_qy REAL, prior_qz REAL, prior_tx REAL,
    prior_ty REAL, prior_tz REAL,
    CONSTRAINT image_id_check CHECK(image_id >= 0 and image_id < 2147483647),
    FOREIGN KEY(camera_id) REFERENCES cameras(camera_id));
CREATE TABLE IF NOT EXISTS keypoints (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
    cols INTEGER NOT NULL, data BLOB, FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS descriptors (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
    cols INTEGER NOT NULL, data BLOB, FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS matches (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
    cols INTEGER NOT NULL, data BLOB);
CREATE TABLE IF NOT EXISTS two_view_geometries (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
    cols INTEGER NOT NULL, data BLOB, config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB, qvec BLOB, tvec BLOB);
CREATE UNIQUE INDEX IF NOT EXISTS index_name ON images(name);
"""


def image_pair_to_pair_id(image_id1, image_id2):
    """R:scene/database.h:28-29."""
    if image_id1 > image_id2:
        image_id1, image_id2 = image_id2, image_id1
    return int(image_id1) * MAX_NUM_IMAGES + int(image_id2)


def pair_id_to_image_pair(pair_id):
    return int(pair_id // MAX_NUM_IMAGES), int(pair_id % MAX_NUM_IMAGES)


class Database:
    def __init__(self, path=None):
        self.con = None
        if path is not None:
            self.open(path)

    @classmethod
    def connect(cls, path):
        return cls(path)

    def open(self, path):
        self.close()
        self.con = sqlite3.connect(os.fspath(path), isolation_level=None)
        self.con.executescript(CREATE_SQL)

    def close(self):
        if self.con is not None:
            self.con.close()
            self.con = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- counts (R:scene/database.h:18-27) ---------------------------------------------------
    def _count(self, table):
        return self.con.execute(f"SELECT COUNT(*) FROM {table}").fetchone()[0]

    num_cameras = property(lambda self: self._count("cameras"))
    num_images = property(lambda self: self._count("images"))
    num_keypoints = property(lambda self: self.con.execute("SELECT COALESCE(SUM(rows),0) FROM keypoints").fetchone()[0])
    num_descriptors = property(
        lambda self: self.con.execute("SELECT COALESCE(SUM(rows),0) FROM descriptors").fetchone()[0])
    num_matches = property(lambda self: self.con.execute("SELECT COALESCE(SUM(rows),0) FROM matches").fetchone()[0])
    num_inlier_matches = property(
        lambda self: self.con.execute("SELECT COALESCE(SUM(rows),0) FROM two_view_geometries").fetchone()[0])
    num_matched_image_pairs = property(
        lambda self: self.con.execute("SELECT COUNT(*) FROM matches WHERE rows > 0").fetchone()[0])
    num_verified_image_pairs = property(
        lambda self: self.con.execute("SELECT COUNT(*) FROM two_view_geometries WHERE rows > 0").fetchone()[0])

    # -- writers used to build databases ----------------------------------------------------------
    def add_camera(self, model, width, height, params, prior_focal_length=False, camera_id=None):
        params = np.asarray(params, np.float64)
        cur = self.con.execute("INSERT INTO cameras VALUES (?, ?, ?, ?, ?, ?)",
                               (camera_id, int(model), int(width), int(height), params.tobytes(),
                                int(bool(prior_focal_length))))
        return cur.lastrowid

    def add_image(self, name, camera_id, image_id=None):
        cur = self.con.execute("INSERT INTO images VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                               (image_id, name, int(camera_id)) + (None,) * 7)
        return cur.lastrowid

    def write_keypoints(self, image_id, keypoints):
        kp = np.ascontiguousarray(keypoints, np.float32)
        assert kp.ndim == 2 and kp.shape[1] in (2, 4, 6)
        self.con.execute("INSERT OR REPLACE INTO keypoints VALUES (?, ?, ?, ?)",
                         (int(image_id), kp.shape[0], kp.shape[1], kp.tobytes()))

    def write_descriptors(self, image_id, descriptors):
        d = np.ascontiguousarray(descriptors, np.uint8)
        assert d.ndim == 2 and d.shape[1] == 128
        self.con.execute("INSERT OR REPLACE INTO descriptors VALUES (?, ?, ?, ?)",
                         (int(image_id), d.shape[0], d.shape[1], d.tobytes()))

    # -- readers -----------------------------------------------------------------------------------
    def read_all_images(self):
        """[(image_id, name, camera_id)] ordered by image_id."""
        return self.con.execute("SELECT image_id, name, camera_id FROM images ORDER BY image_id").fetchall()

    def read_camera(self, camera_id):
        row = self.con.execute("SELECT model, width, height, params, prior_focal_length FROM cameras "
                               "WHERE camera_id = ?", (int(camera_id),)).fetchone()
        if row is None:
            raise ValueError(f"[database.py] Check Failed: camera {camera_id} exists")
        model, w, h, params, prior = row
        if model not in CAMERA_MODEL_NUM_PARAMS:
            raise ValueError(f"[database.py] camera model id {model} is not supported by the B200 verifier "
                             "(COLMAP 3.9.1 model ids 0-10)")
        p = np.frombuffer(params, np.float64)
        if len(p) != CAMERA_MODEL_NUM_PARAMS[model]:
            raise ValueError(f"[database.py] Check Failed: camera model {model} has "
                             f"{CAMERA_MODEL_NUM_PARAMS[model]} parameters")
        return dict(model=int(model), width=int(w), height=int(h), params=[float(x) for x in p],
                    has_prior_focal_length=int(prior))

    def read_keypoints(self, image_id):
        row = self.con.execute("SELECT rows, cols, data FROM keypoints WHERE image_id = ?", (int(image_id),)).fetchone()
        if row is None or row[0] == 0:
            return np.zeros((0, 2), np.float32)
        return np.frombuffer(row[2], np.float32).reshape(row[0], row[1])

    def read_descriptors(self, image_id):
        row = self.con.execute("SELECT rows, cols, data FROM descriptors WHERE image_id = ?", (int(image_id),)).fetchone()
        if row is None or row[0] == 0:
            return np.zeros((0, 128), np.uint8)
        return np.frombuffer(row[2], np.uint8).reshape(row[0], row[1])

    def exists_matches(self, id1, id2):
        return self.con.execute("SELECT 1 FROM matches WHERE pair_id = ?",
                                (image_pair_to_pair_id(id1, id2),)).fetchone() is not None

    def exists_inlier_matches(self, id1, id2):
        return self.con.execute("SELECT 1 FROM two_view_geometries WHERE pair_id = ?",
                                (image_pair_to_pair_id(id1, id2),)).fetchone() is not None

    def existing_pair_ids(self, table):
        return {r[0] for r in self.con.execute(f"SELECT pair_id FROM {table}")}

    def read_matches(self, id1, id2):
        row = self.con.execute("SELECT rows, cols, data FROM matches WHERE pair_id = ?",
                               (image_pair_to_pair_id(id1, id2),)).fetchone()
        if row is None or row[0] == 0:
            return np.zeros((0, 2), np.uint32)
        m = np.frombuffer(row[2], np.uint32).reshape(row[0], 2)
        return m[:, ::-1].copy() if id1 > id2 else m.copy()

    def read_two_view_geometry(self, id1, id2):
        """R:scene/database.h:30-33.  Returns dict(config, F, E, H, inlier_matches)."""
        row = self.con.execute("SELECT rows, cols, data, config, F, E, H, qvec, tvec FROM two_view_geometries WHERE pair_id = ?",
                               (image_pair_to_pair_id(id1, id2),)).fetchone()
        if row is None:
            return None
        rows, cols, data, config, F, E, H, qvec, tvec = row
        inl = np.frombuffer(data, np.uint32).reshape(rows, 2).copy() if rows else np.zeros((0, 2), np.uint32)

        def mat(b):
            return np.frombuffer(b, np.float64).reshape(3, 3).copy() if b else np.zeros((3, 3))
        F, E, H = mat(F), mat(E), mat(H)
        if id1 > id2:  # TwoViewGeometry::Invert
            inl = inl[:, ::-1].copy()
            F, E = F.T.copy(), E.T.copy()
            H = np.linalg.inv(H) if np.abs(H).sum() > 0 else H
        q = np.frombuffer(qvec, np.float64).copy() if qvec and len(qvec) == 32 else np.array([1.0, 0, 0, 0])
        t = np.frombuffer(tvec, np.float64).copy() if tvec and len(tvec) == 24 else np.zeros(3)
        if id1 > id2 and (np.any(t != 0) or not np.array_equal(q, [1.0, 0, 0, 0])):
            w, x, y, z = q
            Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                           [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                           [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            q, t = np.array([w, -x, -y, -z]), -Rm.T @ t
        return dict(config=int(config), F=F, E=E, H=H, inlier_matches=inl, qvec=q, tvec=t)

    # -- result writers (U:scene/database.cc WriteMatches / WriteTwoViewGeometry) --------------------
    def write_matches(self, id1, id2, matches):
        m = np.ascontiguousarray(matches, np.uint32).reshape(-1, 2)
        if id1 > id2:
            m = np.ascontiguousarray(m[:, ::-1])
        self.con.execute("INSERT OR REPLACE INTO matches VALUES (?, ?, ?, ?)",
                         (image_pair_to_pair_id(id1, id2), m.shape[0], 2, m.tobytes()))

    def write_two_view_geometry(self, id1, id2, config, inlier_matches, F=None, E=None, H=None, qvec=None, tvec=None):
        m = np.ascontiguousarray(inlier_matches, np.uint32).reshape(-1, 2)
        F = np.zeros((3, 3)) if F is None else np.asarray(F, np.float64)
        E = np.zeros((3, 3)) if E is None else np.asarray(E, np.float64)
        H = np.zeros((3, 3)) if H is None else np.asarray(H, np.float64)
        if id1 > id2:  # store in the id1 < id2 frame
            m = np.ascontiguousarray(m[:, ::-1])
            F, E = F.T, E.T
            H = np.linalg.inv(H) if np.abs(H).sum() > 0 else H
        qvec = np.array([1.0, 0, 0, 0]) if qvec is None else np.asarray(qvec, np.float64).reshape(4)   # (w, x, y, z)
        tvec = np.zeros(3) if tvec is None else np.asarray(tvec, np.float64).reshape(3)
        if id1 > id2 and (np.any(tvec != 0) or not np.array_equal(qvec, [1.0, 0, 0, 0])):
            # cam2_from_cam1 of the swapped pair = the inverse pose (the identity stays bit-for-bit the identity)
            w, x, y, z = qvec
            Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                           [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                           [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            qvec, tvec = np.array([w, -x, -y, -z]), -Rm.T @ tvec
        self.con.execute("INSERT OR REPLACE INTO two_view_geometries VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                         (image_pair_to_pair_id(id1, id2), m.shape[0], 2, m.tobytes(), int(config),
                          np.ascontiguousarray(F).tobytes(), np.ascontiguousarray(E).tobytes(),
                          np.ascontiguousarray(H).tobytes(), qvec.tobytes(), tvec.tobytes()))

    def transaction(self):
        return _Transaction(self.con)


class _Transaction:
    """DatabaseTransaction (one per block of pairs upstream)."""

    def __init__(self, con):
        self.con = con

    def __enter__(self):
        self.con.execute("BEGIN")
        return self

    def __exit__(self, et, ev, tb):
        self.con.execute("COMMIT" if et is None else "ROLLBACK")
