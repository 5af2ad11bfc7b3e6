"""stand-in for `colored_glog` (pipeline/pipeline_module.py:2): the handful of calls the reference makes"""
import logging

_log = logging.getLogger("nerf_slam")
info, debug, warn, error = _log.info, _log.debug, _log.warning, _log.error


def log(level, msg):
    _log.log(level, msg)


def check(cond, msg=""):
    if not cond:
        raise AssertionError(msg)
