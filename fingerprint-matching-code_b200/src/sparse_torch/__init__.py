from .csx_matrix import CSRMatrix3d
from .csx_matrix import CSCMatrix3d
from .csx_matrix import CSXMatrix3d
from .csx_matrix import concatenate
from .csx_matrix import dot
