"""Option classes with the behaviour of the reference's `MakeDataclass` (R:helpers.h:40-283):
attribute access, construction from a dict or kwargs, `mergedict`, `todict`, `summary`,
copy / deepcopy, pickle, and implicit dict -> Options conversion at call sites.

Defaults follow the C++ structs the bindings expose:
  SiftMatchingOptions       R:pipeline/match_features.h:71-100   (U:feature/sift.h)
  ExhaustiveMatchingOptions R:pipeline/match_features.h:102-108
  SequentialMatchingOptions R:pipeline/match_features.h:110-152
  RANSACOptions             R:optim/bindings.h:7-27  (Python-constructed defaults, :10-18)
  TwoViewGeometryOptions    R:estimators/two_view_geometry.h:41-65 (its `ransac` member keeps the C++
                            constructor defaults: py::init<>() runs the C++ ctor)
"""
import copy
import enum


class Device(enum.Enum):
    """R:utils.h:9 / R:main.cc:98-106."""
    auto = -1
    cpu = 0
    cuda = 1


class TwoViewGeometryConfiguration(enum.IntEnum):
    """R:estimators/two_view_geometry.h:67-80."""
    UNDEFINED = 0
    DEGENERATE = 1
    CALIBRATED = 2
    UNCALIBRATED = 3
    PLANAR = 4
    PANORAMIC = 5
    PLANAR_OR_PANORAMIC = 6
    WATERMARK = 7
    MULTIPLE = 8


def _enum_from(cls, value):
    """AddStringToEnumConstructor (R:helpers.h:45-51): enums are constructible from their name."""
    if isinstance(value, cls):
        return value
    if isinstance(value, str):
        try:
            return cls[value]
        except KeyError:
            raise ValueError(f"Invalid string value {value} for enum {cls.__name__}") from None
    return cls(value)


class _Options:
    """Base of every option class.  Subclasses declare `_fields = {name: (default, doc)}`."""
    _fields = {}

    def __init__(self, *args, **kwargs):
        for k, (default, _doc) in self._fields.items():
            object.__setattr__(self, k, copy.deepcopy(default() if callable(default) else default))
        if len(args) > 1:
            raise TypeError(f"{type(self).__name__}() takes at most one positional argument (a dict)")
        if args:
            if isinstance(args[0], type(self)):
                self.mergedict(args[0].todict())
            elif isinstance(args[0], dict):
                self.mergedict(args[0])
            else:
                raise TypeError(f"{type(self).__name__}(): expected a dict, got {type(args[0]).__name__}")
        if kwargs:
            self.mergedict(kwargs)

    def __setattr__(self, name, value):
        if name not in self._fields:
            raise AttributeError(f"{type(self).__name__} has no attribute '{name}'")
        cur = getattr(self, name)
        object.__setattr__(self, name, self._coerce(name, cur, value))

    def _coerce(self, name, cur, value):
        cls = type(self).__name__
        if isinstance(cur, _Options):
            if isinstance(value, dict):
                new = copy.deepcopy(cur)
                new.mergedict(value)
                return new
            if isinstance(value, type(cur)):
                return value
            raise TypeError(f"{cls}.{name}: expected {type(cur).__name__} or dict, got {type(value).__name__}")
        if isinstance(cur, bool):
            if isinstance(value, (bool, int)) and not isinstance(value, float):
                return bool(value)
        elif isinstance(cur, int):
            if isinstance(value, int) and not isinstance(value, bool):
                return int(value)
            if isinstance(value, bool):
                return int(value)
        elif isinstance(cur, float):
            if isinstance(value, (int, float)) and not isinstance(value, bool):
                return float(value)
        elif isinstance(cur, str):
            if isinstance(value, (str, bytes)) or hasattr(value, "__fspath__"):
                return str(value if not hasattr(value, "__fspath__") else value.__fspath__())
        # readable TypeError like R:helpers.h:87-121
        raise TypeError(f"{cls}.{name}: Could not convert {value!r}: {type(value).__name__} to "
                        f"'{type(cur).__name__}'.")

    def mergedict(self, d):
        """Recursive update from a dict (R:helpers.h:53-124)."""
        if not isinstance(d, dict):
            raise TypeError("mergedict() expects a dict")
        for k, v in d.items():
            if k not in self._fields:
                raise AttributeError(f"{type(self).__name__} has no attribute '{k}'")  # R:helpers.h:62-66
            setattr(self, k, v)

    def todict(self, recursive=True):
        out = {}
        for k in self._fields:
            v = getattr(self, k)
            out[k] = v.todict() if (recursive and isinstance(v, _Options)) else v
        return out

    def summary(self, write_type=False):
        lines = [f"{type(self).__name__}:"]
        for k in self._fields:
            v = getattr(self, k)
            if isinstance(v, _Options):
                sub = v.summary(write_type).split("\n")
                lines.append(f"    {k}: " + sub[0])
                lines += ["    " + s for s in sub[1:]]
            else:
                lines.append(f"    {k}" + (f": {type(v).__name__}" if write_type else "") + f" = {v}")
        return "\n".join(lines)

    def __repr__(self):
        return self.summary()

    def __eq__(self, other):
        return type(other) is type(self) and self.todict() == other.todict()

    def __copy__(self):
        return type(self)(self.todict())

    def __deepcopy__(self, memo):
        return type(self)(copy.deepcopy(self.todict(), memo))

    def __getstate__(self):
        return self.todict()

    def __setstate__(self, state):
        type(self).__init__(self)
        self.mergedict(state)

    @classmethod
    def coerce(cls, value):
        """Implicit dict -> Options at call sites (R:helpers.h:258-268)."""
        if value is None:
            return cls()
        if isinstance(value, cls):
            return value
        if isinstance(value, dict):
            return cls(value)
        raise TypeError(f"expected {cls.__name__} or dict, got {type(value).__name__}")


class SiftMatchingOptions(_Options):
    _fields = {
        "num_threads": (-1, ""),
        "gpu_index": ("-1", "Index of the GPU used for feature matching. For multi-GPU matching, you should "
                            "separate multiple GPU indices by comma, e.g., \"0,1,2,3\"."),
        "max_ratio": (0.8, "Maximum distance ratio between first and second best match."),
        "max_distance": (0.7, "Maximum distance to best match."),
        "cross_check": (True, "Whether to enable cross checking in matching."),
        "max_num_matches": (32768, "Maximum number of matches."),
        "guided_matching": (False, "Whether to perform guided matching, if geometric verification succeeds."),
    }


class ExhaustiveMatchingOptions(_Options):
    _fields = {"block_size": (50, "")}


class SequentialMatchingOptions(_Options):
    _fields = {
        "overlap": (10, "Number of overlapping image pairs."),
        "quadratic_overlap": (True, "Whether to match images against their quadratic neighbors."),
        "loop_detection": (False, "Loop detection is invoked every `loop_detection_period` images."),
        "loop_detection_period": (10, ""),
        "loop_detection_num_images": (50, ""),
        "loop_detection_num_nearest_neighbors": (1, ""),
        "loop_detection_num_checks": (256, ""),
        "loop_detection_num_images_after_verification": (0, ""),
        "loop_detection_max_num_features": (-1, ""),
        "vocab_tree_path": ("", ""),
    }


class RANSACOptions(_Options):
    # Python-constructed defaults of the binding (R:optim/bindings.h:10-18)
    _fields = {
        "max_error": (4.0, ""),
        "min_inlier_ratio": (0.01, ""),
        "confidence": (0.9999, ""),
        "dyn_num_trials_multiplier": (3.0, ""),
        "min_num_trials": (1000, ""),
        "max_num_trials": (100000, ""),
    }


def _cpp_ransac_defaults():
    # colmap::RANSACOptions C++ constructor defaults (what TwoViewGeometryOptions().ransac holds)
    return RANSACOptions(max_error=4.0, min_inlier_ratio=0.25, confidence=0.999, dyn_num_trials_multiplier=3.0,
                         min_num_trials=100, max_num_trials=10000)


class TwoViewGeometryOptions(_Options):
    _fields = {
        "min_num_inliers": (15, ""),
        "min_E_F_inlier_ratio": (0.95, ""),
        "max_H_inlier_ratio": (0.8, ""),
        "watermark_min_inlier_ratio": (0.7, ""),
        "watermark_border_size": (0.1, ""),
        "detect_watermark": (True, ""),
        "multiple_ignore_watermark": (True, ""),
        "force_H_use": (False, ""),
        "compute_relative_pose": (False, ""),
        "multiple_models": (False, ""),
        "ransac": (_cpp_ransac_defaults, ""),
    }
