"""tabcorr_b200 -- B200-native (sm_100a CUDA) implementation of TabCorr's prediction hot path.

Drop-in for ``TabCorr.read`` / ``TabCorr.predict`` / ``Interpolator.predict`` / ``database.read`` of
johannesulf/TabCorr, plus batched ``predict_batch`` entry points.  See DESIGN.md.
"""

__version__ = '0.1.0'

from .tabcorr import TabCorr  # noqa: F401,E402
from .interpolator import Interpolator  # noqa: F401,E402
from . import database  # noqa: F401,E402
from .tableset import TableSet  # noqa: F401,E402
from . import sweep  # noqa: F401,E402
from . import distributed  # noqa: F401,E402
from . import synthetic  # noqa: F401,E402
from .multipole import tabcorr_s_mu_to_multipole, tpcf_multipole  # noqa: F401,E402
from . import models  # noqa: F401,E402
from .models import PrebuiltHodModelFactory  # noqa: F401,E402
