"""The 40 pose scenarios of the reference's own test suite (tests/test_pose_optimization.cpp:264-1295), restated as
a table: (name, true position, true (yaw,pitch,roll), guess factor, point outlier proportion or None, plane outlier
proportion or None). The guess is `factor x truth` on every component, exactly as each TEST builds it
(GOOD 0.9, MEDIUM 0.5, BAD 0.1; :35-37). Tolerances are the reference's: +-(1 + POINTS_ERROR) mm per axis and
+-0.1 degree per Euler angle (:230-242)."""
import numpy as np
from scipy.spatial.transform import Rotation

import oracle_lib as ol

D2R = np.pi / 180.0
END = 10.0
YAW, PITCH, ROLL = 45 * D2R, -45 * D2R, 20 * D2R
GOOD, MEDIUM, BAD = 0.9, 0.5, 0.1
POINTS_ERROR = 5.0
PLANE_ERROR = 5.0

P0 = (0.0, 0.0, 0.0)
PE = (END, END, END)
R0 = (0.0, 0.0, 0.0)
RE = (YAW, PITCH, ROLL)

SCENARIOS = [
    # :264-405 points only, rotation + translation
    ("noRotationNoTranslation", P0, R0, 1.0, 0.0, None),
    ("perfectGuess", PE, RE, 1.0, 0.0, None),
    ("rotationTranslationGoodGuess", PE, RE, GOOD, 0.0, None),
    ("rotationTranslationMediumGuess", PE, RE, MEDIUM, 0.0, None),
    ("rotationTranslationBadGuess", PE, RE, BAD, 0.0, None),
    # :409-492 translation only
    ("translationGoodGuess", PE, R0, GOOD, 0.0, None),
    ("translationMediumGuess", PE, R0, MEDIUM, 0.0, None),
    ("translationBadGuess", PE, R0, BAD, 0.0, None),
    # :496-795 rotation only
    ("rotationYawGoodGuess", P0, (YAW, 0, 0), GOOD, 0.0, None),
    ("rotationPitchGoodGuess", P0, (0, PITCH, 0), GOOD, 0.0, None),
    ("rotationRollGoodGuess", P0, (0, 0, ROLL), GOOD, 0.0, None),
    ("rotationGoodGuess", P0, RE, GOOD, 0.0, None),
    ("rotationYawMediumGuess", P0, (YAW, 0, 0), MEDIUM, 0.0, None),
    ("rotationPitchMediumguess", P0, (0, PITCH, 0), MEDIUM, 0.0, None),
    ("rotationRollMediumGuess", P0, (0, 0, ROLL), MEDIUM, 0.0, None),
    ("rotationMediumGuess", P0, RE, MEDIUM, 0.0, None),
    ("rotationYawBadGuess", P0, (YAW, 0, 0), BAD, 0.0, None),
    ("rotationPitchBadGuess", P0, (0, PITCH, 0), BAD, 0.0, None),
    ("rotationRollBadGuess", P0, (0, 0, ROLL), BAD, 0.0, None),
    ("rotationBadGuess", P0, RE, BAD, 0.0, None),
    # :798-896 four planes only
    ("plane4PerfectGuess", PE, RE, 1.0, None, 0.0),
    ("plane4GoodGuess", PE, RE, GOOD, None, 0.0),
    ("plane4MediumGuess", PE, RE, MEDIUM, None, 0.0),
    ("plane4BadGuess", PE, RE, BAD, None, 0.0),
    # :900-1002 planes + points
    ("multiPerfectFirstGuess", PE, RE, 1.0, 0.0, 0.0),
    ("multiGoodFirstGuess", PE, RE, GOOD, 0.0, 0.0),
    ("multiMediumFirstGuess", PE, RE, MEDIUM, 0.0, 0.0),
    ("multiBadFirstGuess", PE, RE, BAD, 0.0, 0.0),
    # :1006-1147 planes with outliers
    ("planesPerfect_10PercentOutliers", PE, RE, 1.0, None, 0.1),
    ("planesBad_10PercentOutliers", PE, RE, BAD, None, 0.1),
    ("planesPerfect_50PercentOutliers", PE, RE, 1.0, None, 0.5),
    ("planesBad_50PercentOutliers", PE, RE, BAD, None, 0.5),
    ("planesPerfect_100PercentOutliers", PE, RE, 1.0, None, 1.0),
    ("planesBad_100PercentOutliers", PE, RE, BAD, None, 1.0),
    # :1151-1295 planes + points with outliers
    ("multiPerfect_10PercentOutliers", PE, RE, 1.0, 0.1, 0.1),
    ("multiBad_10PercentOutliers", PE, RE, BAD, 0.1, 0.1),
    ("multiPerfect_50PercentOutliers", PE, RE, 1.0, 0.5, 0.5),
    ("multiBad_50PercentOutliers", PE, RE, BAD, 0.5, 0.5),
    ("multiPerfect_100PercentOutliers", PE, RE, 1.0, 1.0, 1.0),
    ("multiBad_100PercentOutliers", PE, RE, BAD, 1.0, 1.0),
]
assert len(SCENARIOS) == 40


def pose7(position, ypr):
    return np.concatenate([np.asarray(position, dtype=np.float64), ol.quaternion_from_euler(*ypr)])


def build(scn):
    """-> (true_pose7, guess_pose7, matches)"""
    _, pos, ypr, g, pt_out, pl_out = scn
    truth = pose7(pos, ypr)
    guess = pose7([g * v for v in pos], [g * v for v in ypr])
    feats = ol.ref_test_features(truth, POINTS_ERROR, -1.0 if pt_out is None else pt_out, PLANE_ERROR,
                                 -1.0 if pl_out is None else pl_out)
    return truth, guess, feats


def euler_xyz(q_wxyz):
    """utils::get_euler_angles_from_quaternion (angle_utils.cpp:14-18): Eigen EulerSystemXYZ -> (yaw, pitch, roll)."""
    w, x, y, z = q_wxyz
    a = Rotation.from_quat([x, y, z, w]).as_euler("XYZ")
    return a[2], a[1], a[0]


def angle_distance(a, b):
    d = np.fmod(abs(a - b), 2 * np.pi)  # test_pose_optimization.cpp:210-214
    return min(d, abs(d - 2 * np.pi))


def check_reference_tolerance(truth, est):
    """run_test_optimization's EXPECT_NEAR / EXPECT_LT block (:230-242). Returns an error string or None."""
    for i in range(3):
        if not abs(truth[i] - est[i]) <= 1 + POINTS_ERROR:
            return "position[%d]: %.4f vs %.4f" % (i, est[i], truth[i])
    te, ee = euler_xyz(truth[3:]), euler_xyz(est[3:])
    for name, a, b in zip(("yaw", "pitch", "roll"), te, ee):
        if not angle_distance(a, b) < 0.1 * D2R:
            return "%s: %.5f vs %.5f rad" % (name, b, a)
    return None
