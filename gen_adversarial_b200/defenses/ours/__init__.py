from .models import (NVAEDefenseModel, CelebaIdentityClassifier)  # noqa: F401
