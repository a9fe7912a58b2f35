from .schedule import WarmupCosine, schedule_from_config  # noqa: F401
