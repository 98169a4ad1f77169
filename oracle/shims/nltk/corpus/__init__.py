from shim_backend import GERMAN_STOP_WORDS


class _Stopwords:
    @staticmethod
    def words(language):
        assert language == "german"
        return list(GERMAN_STOP_WORDS)


stopwords = _Stopwords()
