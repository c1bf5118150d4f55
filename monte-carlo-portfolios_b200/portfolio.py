"""Mirror of the reference loader's codec call sites
(maprdb-portfolio/src/main/java/com/mapr/db/data/LoadPortfolios.java:251-260, 597-623, 625-662),
batched: the reference compresses one portfolio per loop iteration, here all portfolios
of a file go to the GPU in one call.
"""
from __future__ import annotations

import json

import numpy as np

from . import _native as N
from .engine import Context

INTEGER_CODEC_NAME = "bpacking+variablebyte"  # LoadPortfolios.java:726
FLOAT_COMPRESSOR_NAME = "snappy"              # LoadPortfolios.java:723


class LoadPortfolios:
    """The subset of LoadPortfolios that touches the codec (the drop-in boundary)."""

    def __init__(self, ctx: Context | None = None, codec_id: int = N.CODEC_BP32_VB):
        from .codec import default_context
        self.ctx = ctx or default_context()
        self.codec_id = codec_id  # LoadPortfolios.java:725: new Composition(new BinaryPacking(), new VariableByte())

    # -- single list, same signatures as the Java methods
    def getCompressedInstruments(self, instruments) -> bytes:  # :608-623
        return self.ctx.get_compressed_instruments(self.codec_id, instruments)

    def getDecompressedInstruments(self, compressedInstruments: bytes, numberOfIntegers: int) -> np.ndarray:  # :625-662
        return self.ctx.get_decompressed_instruments(self.codec_id, compressedInstruments, numberOfIntegers)

    # -- batched
    def compress_all(self, lists) -> list[bytes]:
        lens = np.array([len(x) for x in lists], np.int64)
        off = np.zeros(len(lists) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        vals = np.concatenate([np.asarray(x, np.int64) for x in lists]) if len(lists) and off[-1] else np.zeros(0, np.int64)
        out_off, out = self.ctx.codec_encode(self.codec_id, N.FRAME_APP, vals, off)
        return [bytes(out[out_off[i]: out_off[i + 1]]) for i in range(len(lists))]

    @staticmethod
    def integersToBytes(instruments) -> bytes:  # BitManipulationHelper.java:17-27 (the "none" fallback; pure byte order)
        return np.asarray(instruments, np.int64).astype(np.uint32).astype(">u4").tobytes()

    def getInstrumentDocument(self, instruments, compressed: bytes | None = None) -> dict:
        """LoadPortfolios.java:254-260 + 597-606: store compressed iff smaller than 4n bytes, else raw big-endian."""
        n = len(instruments)
        blob = self.getCompressedInstruments(instruments) if compressed is None else compressed
        algo = INTEGER_CODEC_NAME
        if not (len(blob) < 4 * n):
            blob, algo = self.integersToBytes(instruments), "none"
        return {"uncompressedIntegerSize": n, "compressed": blob, "compressedByteSize": len(blob), "algo": algo}

    # -- the weights column (LoadPortfolios.java:563-595: Snappy.compress(double[]); raw Snappy format, see include/mcp.h)
    def getCompressedWeights(self, weights) -> bytes:
        return self.ctx.get_compressed_weights(weights)

    def getDecompressedWeights(self, compressedWeights: bytes, numberOfDoubles: int) -> np.ndarray:  # Snappy.uncompressDoubleArray
        return self.ctx.get_decompressed_weights(compressedWeights, numberOfDoubles)

    @staticmethod
    def doublesToBytes(weights) -> bytes:  # BitManipulationHelper.java:193-208 (the "none" fallback: big-endian IEEE bits)
        return np.asarray(weights, np.float64).astype(">f8").tobytes()

    def getWeightDocument(self, weights, compressed: bytes | None = None) -> dict:
        """LoadPortfolios.java:270-278 + 553-561: the Snappy blob iff it is smaller than 8 n bytes, else raw big-endian."""
        n = len(weights)
        blob = self.getCompressedWeights(weights) if compressed is None else compressed
        algo = FLOAT_COMPRESSOR_NAME
        if not (len(blob) < 8 * n):
            blob, algo = self.doublesToBytes(weights), "none"
        return {"uncompressedDoubleSize": n, "compressed": blob, "compressedByteSize": len(blob), "algo": algo}

    def documents_from_json(self, path: str) -> list[dict]:
        """One document per JSON line of a portfolios.json-style file (schema: SURVEY.md A.8): instruments and weights."""
        rows = [json.loads(line) for line in open(path) if line.strip()]
        lists = [[int(p["instrument"]) for p in r["instruments"]] for r in rows]
        blobs = self.compress_all(lists)
        wl = [[float(p["weight"]) for p in r["instruments"]] for r in rows]
        woff = np.zeros(len(wl) + 1, np.int64)
        np.cumsum([len(x) for x in wl], out=woff[1:])
        wo, wb = self.ctx.weights_encode(np.concatenate([np.asarray(x, np.float64) for x in wl]) if woff[-1] else np.zeros(0), woff)
        docs = []
        for i, (r, ins, blob, w) in enumerate(zip(rows, lists, blobs, wl)):
            docs.append({"_id": str(r.get("id", len(docs))),
                         "metadata": {"min_instrument": min(ins), "max_instrument": max(ins),
                                      "min_weight": min(w), "max_weight": max(w)},
                         "instruments": self.getInstrumentDocument(ins, blob),
                         "weights": self.getWeightDocument(w, bytes(wb[wo[i]: wo[i + 1]]))})
        return docs


def read_portfolios_json(path: str):
    """portfolios.json (one JSON object per line, LoadPortfolios.java:177-419 / synth/portfolios.synth): -> ids, and the
    positions in CSR form (instruments int32, weights float64, pos_off int64[P+1]) in file order."""
    ids, inst, w, off = [], [], [], [0]
    for line in open(path):
        if not line.strip():
            continue
        r = json.loads(line)
        ids.append(r.get("id", len(ids)))
        for pos in r["instruments"]:
            inst.append(int(pos["instrument"]))
            w.append(float(pos["weight"]))
        off.append(len(inst))
    return ids, np.asarray(inst, np.int32), np.asarray(w, np.float64), np.asarray(off, np.int64)


def read_macro_json(path: str) -> np.ndarray:
    """macro.json (synth/macro.synth): one row per date_offset, factors m0..m{F-1} -> S[N][F] (row = trial)."""
    rows = [json.loads(line) for line in open(path) if line.strip()]
    rows.sort(key=lambda r: r.get("date_offset", 0))
    F = sum(1 for k in rows[0] if k.startswith("m") and k[1:].isdigit())
    return np.array([[float(r[f"m{k}"]) for k in range(F)] for r in rows], np.float64)


RESULTS_TABLE = "/apps/simulation_results"  # created, never written, by the reference: ManageMapRDBTable.java:15
BREACH_CODEC_NAME = "delta+fastpfor+variablebyte"


def simulation_documents(ids, result: dict, alpha: float, algo: str = BREACH_CODEC_NAME) -> list[dict]:
    """One sink document per portfolio from a run's result arrays (Context.simulate / fetch, or sharding.assemble),
    shaped like the loader's instrument document (LoadPortfolios.java:597-606) and with its storage rule (:254-260):
    the encoded breach list is stored only if it is smaller than the raw big-endian ints, else those with algo "none".
    The blob decodes with the stock Java library (Composition(FastPFOR, VariableByte) + Delta.fastinverseDelta)."""
    t = int(result["tail"])
    enc, off = result["enc"], result["enc_off"]
    docs = []
    for p, pid in enumerate(ids):
        blob = bytes(enc[off[p]: off[p + 1]])
        a = algo
        if not (len(blob) < 4 * t):
            if result.get("breach") is None:
                raise ValueError("the raw breach lists are needed for the uncompressed fallback")
            blob, a = LoadPortfolios.integersToBytes(result["breach"][p]), "none"
        docs.append({"_id": str(pid), "alpha": alpha, "trials": int(result.get("trials", 0)),
                     "mean": float(result["mean"][p]), "variance": float(result["variance"][p]),
                     "var": {"loss": float(result["var_loss"][p]), "trial": int(result["var_trial"][p])},
                     "cvar": float(result["cvar"][p]),
                     "breaches": {"uncompressedIntegerSize": t, "compressed": blob, "compressedByteSize": len(blob), "algo": a}})
    return docs


class SimulatePortfolios:
    """The entry point the reference names a table for but never wrote: portfolios.json x macro.json -> exposures on the
    device (mcp_build_exposures) -> the hot path -> one document per portfolio for /apps/simulation_results."""

    def __init__(self, ctx: Context | None = None, codec_id: int = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA):
        from .codec import default_context
        self.ctx = ctx or default_context()
        self.codec_id = codec_id

    def run(self, portfolios_json: str, macro_json: str, loadings, alpha: float) -> list[dict]:
        ids, inst, w, off = read_portfolios_json(portfolios_json)
        S = read_macro_json(macro_json)
        self.ctx.build_exposures(inst, w, off, loadings)
        self.ctx.upload_scenarios(S)
        self.ctx.run(alpha, self.codec_id, N.FRAME_APP)
        r = self.ctx.fetch()
        r["trials"] = S.shape[0]
        return simulation_documents(ids, r, alpha)
