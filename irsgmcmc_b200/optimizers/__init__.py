from .adam_rate_decay import Adam
