package ch.sqooba.kao;

import java.util.*;

/**
 * Host-side model builder for KaoNative.solve — the place where the reference builds its LP
 * (README.md:139-185).  Same defaults as docs/MODEL.md §1 / kao_cli.cpp.  UNCOMPILED here.
 * JSON parsing is left to the tool's existing JSON library (the snapshot does not show which).
 */
public final class AssignmentSolver {
    /** current[p] = broker ids, leader first (README.md:52-63); brokers = target list (README.md:48). */
    public static int[][] solve(int[][] current, int[] brokers, Map<Integer, String> rackOfBroker, int rf,
                                long seed, int rounds, int roundSize, int device) {
        int[] ids = Arrays.stream(brokers).sorted().distinct().toArray();
        int P = current.length, B = ids.length;
        Map<Integer, Integer> dense = new HashMap<>();
        for (int i = 0; i < B; i++) dense.put(ids[i], i);
        List<String> racks = new ArrayList<>(new TreeSet<>(rackOfBroker.values()));
        racks.removeIf(r -> Arrays.stream(ids).noneMatch(b -> rackOfBroker.get(b).equals(r)));
        int R = racks.size();
        byte[] rackOf = new byte[B];
        long[] size = new long[R];
        for (int i = 0; i < B; i++) { rackOf[i] = (byte) racks.indexOf(rackOfBroker.get(ids[i])); size[rackOf[i]]++; }
        int rfCur = Math.max(1, Arrays.stream(current).mapToInt(c -> c.length).max().orElse(1));
        int[] cur = new int[P * rfCur];
        Arrays.fill(cur, -1);
        short[] wF = new short[P * B], wL = new short[P * B];
        int[] WL = {4, 2, 1}, WF = {2, 2, 1};                       // README.md:146, :131-133
        for (int p = 0; p < P; p++)
            for (int i = 0; i < current[p].length; i++) {
                Integer d = dense.get(current[p][i]);
                if (d == null) continue;                           // broker leaves the cluster
                cur[p * rfCur + i] = d;
                wF[p * B + d] = (short) (i < 3 ? WF[i] : 1);
                wL[p * B + d] = (short) (i < 3 ? WL[i] : 1);
            }
        long tot = (long) P * rf;
        int[] bounds = new int[4 * B + 2 * R + 2];
        for (int b = 0; b < B; b++) {
            bounds[b] = (int) (tot / B);         bounds[B + b] = (int) ((tot + B - 1) / B);       // C3
            bounds[2 * B + b] = P / B;           bounds[3 * B + b] = (P + B - 1) / B;             // C4
        }
        for (int r = 0; r < R; r++) {                                                           // C6
            bounds[4 * B + r] = (int) (tot * size[r] / B);
            bounds[4 * B + R + r] = (int) ((tot * size[r] + B - 1) / B);
        }
        bounds[4 * B + 2 * R] = rf / R;  bounds[4 * B + 2 * R + 1] = (rf + R - 1) / R;            // C7
        int[] out = new int[P * rf];
        long[] stats = new long[4];
        KaoNative.solve(P, B, R, rf, rfCur, rackOf, wF, wL, bounds, cur, seed, rounds, roundSize, device, out, stats);
        int[][] res = new int[P][];
        for (int p = 0; p < P; p++) {
            int n = 0;
            for (int i = 0; i < rf; i++) if (out[p * rf + i] >= 0) n++;
            res[p] = new int[n];
            for (int i = 0, k = 0; i < rf; i++) if (out[p * rf + i] >= 0) res[p][k++] = ids[out[p * rf + i]];
        }
        return res;                                                // leader first (README.md:67-78, :88)
    }
}
