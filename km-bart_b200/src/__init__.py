"""`src` is also the name of the reference's package (src.data, src.training, src.generation, ...).
Merging the two as a namespace lets the reference's unchanged scripts run with this directory placed
BEFORE the KM-BART checkout on PYTHONPATH: `src.model` resolves here (first on __path__), every other
`src.*` subpackage resolves to the checkout."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
