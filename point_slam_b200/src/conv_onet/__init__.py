from . import config, models
