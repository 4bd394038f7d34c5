"""Stand-in for nose_parameterized.parameterized used at
/root/reference/bayesic/tests/test_algebra.py:13,208,216,234 -- loops the cases."""


def parameterized(cases):
    def deco(fn):
        def run_all():
            for case in cases:
                fn(*case)
        run_all.__name__ = fn.__name__
        return run_all
    return deco
