"""Stub of the `colorama` package (not installed here): the reference imports it for table colours
only (R/bayes_opt/target_space.py:10, R/bayes_opt/logger.py:8).  Used by make_golden.py only."""


class _Codes:
    def __getattr__(self, name):
        return ""


Fore = _Codes()
Back = _Codes()
Style = _Codes()


def just_fix_windows_console():
    pass


def init(*args, **kwargs):
    pass
