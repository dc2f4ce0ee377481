"""COLMAP SQLite database access for the data either side of the hot path (keypoints in, matches in, refined
keypoints out).  Own implementation of the COLMAP 3.x database *format* (the schema and blob encodings documented
by COLMAP: float32 keypoint rows, uint8 descriptors, uint32 match pairs, pair_id = id_small * (2^31 - 1) + id_large);
the reference ships COLMAP's scripts/python/database.py as `pixsfm/util/database.py` and uses it from
`pixsfm/util/colmap.py:9-69`.  Only the Python standard library (sqlite3) and numpy are needed."""
import sqlite3

import numpy as np

MAX_IMAGE_ID = 2 ** 31 - 1

_SCHEMA = (
    """CREATE TABLE IF NOT EXISTS cameras (
        camera_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, model INTEGER NOT NULL, width INTEGER NOT NULL,
        height INTEGER NOT NULL, params BLOB, prior_focal_length INTEGER NOT NULL)""",
    """CREATE TABLE IF NOT EXISTS images (
        image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE, camera_id INTEGER NOT NULL,
        prior_qw REAL, prior_qx REAL, prior_qy REAL, prior_qz REAL, prior_tx REAL, prior_ty REAL, prior_tz REAL,
        CONSTRAINT image_id_check CHECK(image_id >= 0 and image_id < %d),
        FOREIGN KEY(camera_id) REFERENCES cameras(camera_id))""" % MAX_IMAGE_ID,
    """CREATE TABLE IF NOT EXISTS keypoints (
        image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
        FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE)""",
    """CREATE TABLE IF NOT EXISTS descriptors (
        image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
        FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE)""",
    """CREATE TABLE IF NOT EXISTS matches (
        pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB)""",
    """CREATE TABLE IF NOT EXISTS two_view_geometries (
        pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
        config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB)""",
    "CREATE UNIQUE INDEX IF NOT EXISTS index_name ON images(name)",
)


def image_ids_to_pair_id(image_id1, image_id2):
    lo, hi = (image_id1, image_id2) if image_id1 <= image_id2 else (image_id2, image_id1)
    return int(lo) * MAX_IMAGE_ID + int(hi)


def pair_id_to_image_ids(pair_id):
    lo, hi = divmod(int(pair_id), MAX_IMAGE_ID)
    return lo, hi


def array_to_blob(array):
    return np.ascontiguousarray(array).tobytes()


def blob_to_array(blob, dtype, shape=(-1,)):
    return np.frombuffer(blob, dtype=dtype).reshape(*shape).copy()


class COLMAPDatabase(sqlite3.Connection):
    @staticmethod
    def connect(database_path):
        return sqlite3.connect(str(database_path), factory=COLMAPDatabase)

    def create_tables(self):
        for stmt in _SCHEMA:
            self.execute(stmt)
        self.commit()

    def add_camera(self, model, width, height, params, prior_focal_length=False, camera_id=None):
        cur = self.execute("INSERT INTO cameras VALUES (?, ?, ?, ?, ?, ?)",
                           (camera_id, int(model), int(width), int(height),
                            array_to_blob(np.asarray(params, np.float64)), int(bool(prior_focal_length))))
        return cur.lastrowid

    def add_image(self, name, camera_id, prior_q=(1.0, 0.0, 0.0, 0.0), prior_t=(0.0, 0.0, 0.0), image_id=None):
        cur = self.execute("INSERT INTO images VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                           (image_id, name, int(camera_id)) + tuple(float(v) for v in prior_q)
                           + tuple(float(v) for v in prior_t))
        return cur.lastrowid

    def add_keypoints(self, image_id, keypoints):
        kp = np.asarray(keypoints, np.float32)
        if kp.ndim != 2 or kp.shape[1] not in (2, 4, 6):
            raise ValueError("keypoints must be [N,2], [N,4] or [N,6]")
        self.execute("INSERT INTO keypoints VALUES (?, ?, ?, ?)", (int(image_id), kp.shape[0], kp.shape[1], array_to_blob(kp)))

    def add_descriptors(self, image_id, descriptors):
        d = np.ascontiguousarray(descriptors, np.uint8)
        self.execute("INSERT INTO descriptors VALUES (?, ?, ?, ?)", (int(image_id), d.shape[0], d.shape[1], array_to_blob(d)))

    def add_matches(self, image_id1, image_id2, matches):
        m = np.asarray(matches, np.uint32)
        if m.ndim != 2 or m.shape[1] != 2:
            raise ValueError("matches must be [N,2]")
        if image_id1 > image_id2:       # stored in the order of the pair id (smaller image id first)
            m = m[:, ::-1]
        self.execute("INSERT INTO matches VALUES (?, ?, ?, ?)",
                     (image_ids_to_pair_id(image_id1, image_id2), m.shape[0], m.shape[1], array_to_blob(m)))

    def add_two_view_geometry(self, image_id1, image_id2, matches, F=None, E=None, H=None, config=2):
        m = np.asarray(matches, np.uint32)
        if m.ndim != 2 or m.shape[1] != 2:
            raise ValueError("matches must be [N,2]")
        if image_id1 > image_id2:
            m = m[:, ::-1]
        mats = [array_to_blob(np.asarray(np.eye(3) if x is None else x, np.float64)) for x in (F, E, H)]
        self.execute("INSERT INTO two_view_geometries VALUES (?, ?, ?, ?, ?, ?, ?, ?)",
                     (image_ids_to_pair_id(image_id1, image_id2), m.shape[0], m.shape[1], array_to_blob(m), int(config), *mats))

    def image_id_to_name(self):
        return {int(i): n for i, n in self.execute("SELECT image_id, name FROM images")}
