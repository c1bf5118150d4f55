package com.mapr.db.data;

import java.io.BufferedReader;
import java.io.File;
import java.io.FileReader;
import java.io.IOException;
import java.util.ArrayList;
import java.util.List;

import com.fasterxml.jackson.core.JsonFactory;
import com.fasterxml.jackson.core.JsonParser;
import com.fasterxml.jackson.core.JsonToken;

import org.ojai.Document;

import com.mapr.db.MapRDB;
import com.mapr.db.Table;

/**
 * The entry point the reference lacks: it ships portfolios.json, macro.json and a table /apps/simulation_results
 * (ManageMapRDBTable.java:15) that nothing writes.  This main values every portfolio under every macro scenario on the
 * GPU (ONE native call: mcp_simulate), and stores one document per portfolio shaped like the loader's instrument document
 * (LoadPortfolios.java:597-606):
 *
 * <pre>
 * { _id, mean, variance, var, var_trial, cvar,
 *   breaches: { uncompressedIntegerSize, compressed (BINARY), compressedByteSize, algo } }
 * </pre>
 *
 * The blob decodes with the stock library: new Composition(new FastPFOR(), new VariableByte()).uncompress(...) after
 * BitManipulationHelper.bytesToInts, then Delta.fastinverseDelta -- exactly how the loader reads its own blobs
 * (LoadPortfolios.java:625-662).
 *
 * Usage: SimulatePortfolios portfolios.json macro.json loadings.csv [alpha]
 *   loadings.csv: one row per instrument id, F comma-separated factor loadings (the reference defines no
 *   instrument-to-factor model: DESIGN.md 3.0).  Exposure of portfolio p to factor f = sum over its positions, in file
 *   order, of weight * loadings[instrument][f].
 */
public class SimulatePortfolios {

    public static final String SIM_RES_TABLE_PATH = "/apps/simulation_results";
    public static final String ALGO = "delta+fastpfor+variablebyte";

    public static void main(String[] args) throws IOException {
        if (args.length < 3) {
            System.err.println("usage: SimulatePortfolios portfolios.json macro.json loadings.csv [alpha]");
            System.exit(2);
        }
        double alpha = args.length > 3 ? Double.parseDouble(args[3]) : 0.99;
        new SimulatePortfolios().run(new File(args[0]), new File(args[1]), new File(args[2]), alpha);
    }

    private void run(File portfolios, File macro, File loadingsFile, double alpha) throws IOException {
        double[][] loadings = readLoadings(loadingsFile);
        int F = loadings[0].length;
        List<String> ids = new ArrayList<String>();
        double[] E = readExposures(portfolios, loadings, F, ids);
        double[] S = readScenarios(macro, F);
        long P = ids.size();
        long N = S.length / F;

        long ctx = McpNative.create(0);
        try {
            double[] stats = new double[(int) (4 * P)];
            int[] varTrial = new int[(int) P];
            long[] encOff = new long[(int) P + 1];
            byte[] enc = McpNative.simulate(ctx, E, P, S, N, F, alpha, McpNative.FASTPFOR256_VB | McpNative.FLAG_DELTA, stats,
                    varTrial, encOff);
            int tail = (int) McpNative.tailSize(N, alpha);

            Table table = MapRDB.getTable(SIM_RES_TABLE_PATH);
            for (int p = 0; p < P; p++) {
                int len = (int) (encOff[p + 1] - encOff[p]);
                byte[] blob = new byte[len];
                System.arraycopy(enc, (int) encOff[p], blob, 0, len);
                Document breaches = MapRDB.newDocument();
                /* the loader stores a blob only if it is smaller than the raw ints (LoadPortfolios.java:254-260) */
                breaches.set("uncompressedIntegerSize", tail)
                        .set("compressed", blob)
                        .set("compressedByteSize", blob.length)
                        .set("algo", ALGO);
                Document doc = MapRDB.newDocument();
                doc.set("_id", ids.get(p))
                        .set("mean", stats[p])
                        .set("variance", stats[(int) P + p])
                        .set("var", stats[2 * (int) P + p])
                        .set("var_trial", varTrial[p])
                        .set("cvar", stats[3 * (int) P + p])
                        .set("breaches", breaches);
                table.insertOrReplace(doc);
            }
            table.flush();
            table.close();
        } finally {
            McpNative.destroy(ctx);
        }
    }

    private static double[][] readLoadings(File f) throws IOException {
        List<double[]> rows = new ArrayList<double[]>();
        BufferedReader r = new BufferedReader(new FileReader(f));
        try {
            String line;
            while ((line = r.readLine()) != null) {
                line = line.trim();
                if (line.length() == 0) {
                    continue;
                }
                String[] parts = line.split(",");
                double[] row = new double[parts.length];
                for (int i = 0; i < parts.length; i++) {
                    row[i] = Double.parseDouble(parts[i]);
                }
                rows.add(row);
            }
        } finally {
            r.close();
        }
        return rows.toArray(new double[rows.size()][]);
    }

    /** portfolios.json: JSON lines {"id":.., "instruments":[{"instrument":i,"weight":w},..]} (synth/portfolios.synth). */
    private static double[] readExposures(File f, double[][] loadings, int F, List<String> ids) throws IOException {
        List<double[]> rows = new ArrayList<double[]>();
        JsonParser jp = new JsonFactory().createParser(f);
        try {
            double[] row = null;
            int instrument = -1;
            String field = null;
            int depth = 0;
            for (JsonToken t = jp.nextToken(); t != null; t = jp.nextToken()) {
                if (t == JsonToken.START_OBJECT) {
                    depth++;
                    if (depth == 1) {
                        row = new double[F];
                    }
                } else if (t == JsonToken.END_OBJECT) {
                    depth--;
                    if (depth == 0) {
                        rows.add(row);
                    }
                } else if (t == JsonToken.FIELD_NAME) {
                    field = jp.getCurrentName();
                } else if (t.isScalarValue()) {
                    if (depth == 1 && "id".equals(field)) {
                        ids.add(jp.getText());
                    } else if (depth == 2 && "instrument".equals(field)) {
                        instrument = jp.getIntValue();
                    } else if (depth == 2 && "weight".equals(field)) {
                        double w = jp.getDoubleValue();
                        for (int k = 0; k < F; k++) {
                            row[k] = Math.fma(w, loadings[instrument][k], row[k]);
                        }
                    }
                }
            }
        } finally {
            jp.close();
        }
        double[] E = new double[rows.size() * F];
        for (int p = 0; p < rows.size(); p++) {
            System.arraycopy(rows.get(p), 0, E, p * F, F);
        }
        return E;
    }

    /** macro.json: JSON lines with keys m0..m{F-1}; row i = trial i (synth/macro.synth). */
    private static double[] readScenarios(File f, int F) throws IOException {
        List<double[]> rows = new ArrayList<double[]>();
        JsonParser jp = new JsonFactory().createParser(f);
        try {
            double[] row = null;
            String field = null;
            for (JsonToken t = jp.nextToken(); t != null; t = jp.nextToken()) {
                if (t == JsonToken.START_OBJECT) {
                    row = new double[F];
                } else if (t == JsonToken.END_OBJECT) {
                    rows.add(row);
                } else if (t == JsonToken.FIELD_NAME) {
                    field = jp.getCurrentName();
                } else if (t.isNumeric() && field != null && field.startsWith("m")) {
                    int k = Integer.parseInt(field.substring(1));
                    if (k < F) {
                        row[k] = jp.getDoubleValue();
                    }
                }
            }
        } finally {
            jp.close();
        }
        double[] S = new double[rows.size() * F];
        for (int n = 0; n < rows.size(); n++) {
            System.arraycopy(rows.get(n), 0, S, n * F, F);
        }
        return S;
    }
}
