LOG_FORMAT = "%(asctime)s\t%(levelname)s\t%(name)s\t%(message)s"

HAP, POP, SUEP = "hap", "pop", "suep"
COHORTS = [HAP, POP, SUEP]
