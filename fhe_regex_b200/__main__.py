"""`python -m fhe_regex_b200 <content> <pattern>` -- the reference's demo (src/main.rs:9-23, src/regex/mod.rs:9-19)
on the B200 backend: gen keys -> encrypt_str -> has_match -> decrypt -> `res: 0|1`."""
import logging
import os
import sys

from . import ClientKey, ServerKey, encrypt_str, has_match, parse


def main(argv):
    if len(argv) != 3:
        print("usage: python -m fhe_regex_b200 <content> <pattern>", file=sys.stderr)
        return 2
    logging.basicConfig(level=os.environ.get("RUST_LOG", "info").upper())      # env_logger default (main.rs:10-11)
    content, pattern = argv[1], argv[2]
    logging.info("parsed: %s", parse(pattern))                                  # main.rs:17-20 (raises on a parse error)
    here = os.path.dirname(os.path.abspath(__file__))
    ck = ClientKey.load(os.path.join(here, "..", "tests", "golden", "client_key"))   # fixture secret key instead of fresh keygen
    sk = ServerKey(keygen_from=ck, seed=0)                                      # ServerKey::new(&client_key), on the GPU
    logging.info("encrypting content..")
    ct = encrypt_str(ck, content)
    logging.info("applying regex..")
    res, st = has_match(sk, ct, pattern, return_stats=True)
    logging.info("%d ciphertext operations, %d cache hits", st["ct_ops"], st["cache_hits"])   # engine.rs:36-40
    print("res: %d" % ck.decrypt(res))                                          # mod.rs:18
    sk.close()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
