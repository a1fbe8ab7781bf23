"""pytest plugin (``-p brian2_b200.pytest_plugin``): registers the ``b200`` device and its
preferences in every pytest process, including pytest-xdist workers, so that the reference's
own test-suite can run with ``test_standalone='b200'`` (tests/tools/run_reference_suite.py)."""
import brian2_b200  # noqa: F401
