"""Drop-in `cmf` package: same import paths as the reference (`from cmf.models import get_model`)."""
