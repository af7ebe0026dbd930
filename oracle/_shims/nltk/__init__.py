def sent_tokenize(text):
    return [text]
