"""Names shared across the package: the three NAPKON cohorts and the log line layout
(tab separated: time, level, logger, message)."""

COHORTS = ["hap", "pop", "suep"]
HAP, POP, SUEP = COHORTS

LOG_FORMAT = "\t".join("%({})s".format(field) for field in ("asctime", "levelname", "name", "message"))
