"""The OpenCV bridge lives with the rest of the test infrastructure in oracle/cv_bridge.py; tests import it from here."""
from oracle.cv_bridge import BRIDGE, Bridge, BridgeMat  # noqa: F401
