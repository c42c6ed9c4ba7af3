"""Minimal `tensorflow` stand-in so the reference's TF-free host modules
(data_utils, tokenizer, text_encoder) import under py3.12 in the build
container.  TEST INFRASTRUCTURE ONLY -- used by tests/golden/make_golden.py to
run the real reference ranking / parsing code and freeze its outputs as
fixtures.  Provides only gfile, logging, compat.as_str."""
import glob as _glob
import logging as _logging
import os as _os


class _GFile:
    @staticmethod
    def Open(name, mode="r"):
        return open(name, mode)

    GFile = Open

    @staticmethod
    def Glob(pattern):
        return _glob.glob(pattern)

    @staticmethod
    def Exists(path):
        return _os.path.exists(path)


gfile = _GFile()


class _Logging:
    info = staticmethod(_logging.info)
    warning = staticmethod(_logging.warning)
    warn = staticmethod(_logging.warning)
    error = staticmethod(_logging.error)


logging = _Logging()


class _Compat:
    @staticmethod
    def as_str(s):
        return s.decode("utf-8") if isinstance(s, bytes) else str(s)

    as_text = as_str


compat = _Compat()
