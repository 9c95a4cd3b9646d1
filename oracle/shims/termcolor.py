"""Shim -- `termcolor.colored` for the verbatim reference param_test_env.py (PaintRLEnv/param_test_env.py:5).
TEST INFRASTRUCTURE ONLY."""


def colored(text, *args, **kwargs):
    return str(text)
