// ZpaqB200.cs -- P/Invoke binding of libzpaqb200.so for ZPAQSharp (source only: this image has no
// .NET toolchain, so the file is not compiled or tested here; the same entry points are exercised
// through ctypes in zpaqsharp_b200/libzpaq.py and tests/).
//
// Drop-in: LibZPAQ.Compress / compressBlock / decompress (LibZPAQ.cs:84,117,65) keep their
// signatures; their bodies gather blocks from the Reader, call the batch ABI once, and write the
// result to the Writer.  LibZPAQ.error (LibZPAQ.cs:22-24) keeps its role: a negative return code
// becomes error(zpq_last_error()).
using System;
using System.Runtime.InteropServices;
using System.Text;

namespace ZPAQSharp
{
    internal static unsafe class ZpaqB200Native
    {
        const string Lib = "zpaqb200";   // libzpaqb200.so / zpaqb200.dll

        [DllImport(Lib)] public static extern int zpq_create(int* deviceIds, int ndev, out IntPtr ctx);
        [DllImport(Lib)] public static extern void zpq_destroy(IntPtr ctx);
        [DllImport(Lib)] public static extern IntPtr zpq_last_error(IntPtr ctx);
        [DllImport(Lib)] public static extern int zpq_compress_blocks(IntPtr ctx, byte* input, ulong* inOff, uint nb,
            [MarshalAs(UnmanagedType.LPStr)] string method, [MarshalAs(UnmanagedType.LPStr)] string filename0,
            [MarshalAs(UnmanagedType.LPStr)] string comment0, int dosha1, byte* output, ulong outCap, ulong* outOff);
        [DllImport(Lib)] public static extern int zpq_compress_blocks_level(IntPtr ctx, int level, byte* input, ulong* inOff, uint nb,
            [MarshalAs(UnmanagedType.LPStr)] string filename0, [MarshalAs(UnmanagedType.LPStr)] string comment0,
            int dosha1, int withTag, byte* output, ulong outCap, ulong* outOff);
        [DllImport(Lib)] public static extern int zpq_compress_blocks_model(IntPtr ctx, byte* hdr, ulong hdrLen, byte* pcomp, ulong pcompLen,
            int* args9, byte* input, ulong* inOff, uint nb, [MarshalAs(UnmanagedType.LPStr)] string filename0,
            [MarshalAs(UnmanagedType.LPStr)] string comment0, int dosha1, int withTag, byte* output, ulong outCap, ulong* outOff);
        [DllImport(Lib)] public static extern long zpq_find_blocks(byte* archive, ulong n, ulong* offsets, ulong maxBlocks);
        [DllImport(Lib)] public static extern long zpq_decompressed_bound(byte* input, ulong* inOff, uint nb);
        [DllImport(Lib)] public static extern int zpq_decompress_blocks(IntPtr ctx, byte* input, ulong* inOff, uint nb, byte* output,
            ulong outCap, ulong* outOff, byte* sha1Status, byte* blockStatus);
        [DllImport(Lib)] public static extern long zpq_make_config([MarshalAs(UnmanagedType.LPStr)] string method, int* args9,
            byte* text, ulong textCap);
        [DllImport(Lib)] public static extern int zpq_compile_config([MarshalAs(UnmanagedType.LPStr)] string config, int* args9,
            byte* hdr, ulong hdrCap, out ulong hdrLen, byte* pcomp, ulong pcompCap, out ulong pcompLen);
        [DllImport(Lib)] public static extern long zpq_builtin_model(int level, byte* hdr, ulong hdrCap);
        [DllImport(Lib)] public static extern double zpq_block_memory(byte* hdr, ulong hdrLen);
        [DllImport(Lib)] public static extern long zpq_device_state_bytes(byte* hdr, ulong hdrLen, int forDecode);
        [DllImport(Lib)] public static extern int zpq_encoder_plan(byte* hdr, ulong hdrLen, uint smemBytes, uint blocksPerSm, int* out8);
    }

    /// <summary>GPU-backed bodies for the LibZPAQ entry points of the compress/decompress path.</summary>
    public static unsafe class LibZPAQB200
    {
        static IntPtr ctx;

        static IntPtr Ctx()
        {
            if (ctx == IntPtr.Zero && ZpaqB200Native.zpq_create(null, 0, out ctx) != 0)
                LibZPAQ.error(Marshal.PtrToStringAnsi(ZpaqB200Native.zpq_last_error(IntPtr.Zero)));
            return ctx;
        }

        static void Check(int rc)
        {
            if (rc != 0) LibZPAQ.error(Marshal.PtrToStringAnsi(ZpaqB200Native.zpq_last_error(ctx)));
        }

        /// <summary>LibZPAQ.Compress (LibZPAQ.cs:84-108): split by the method's block size, code all blocks at once.</summary>
        public static void Compress(Reader input, Writer output, string method, string filename = null, string comment = null, bool dosha1 = true)
        {
            int bs = 4;
            if (method.Length > 1 && char.IsDigit(method[1]))
            {
                bs = method[1] - '0';
                if (method.Length > 2 && char.IsDigit(method[2])) bs = bs * 10 + method[2] - '0';
                if (bs > 11) bs = 11;
            }
            long blockSize = (0x100000L << bs) - 4096;
            var data = new System.IO.MemoryStream();
            var buf = new char[1 << 16];
            int n;
            while ((n = input.read(buf, buf.Length)) > 0)
                for (int i = 0; i < n; ++i) data.WriteByte((byte)buf[i]);
            byte[] all = data.ToArray();
            uint nb = (uint)((all.Length + blockSize - 1) / blockSize);
            if (nb == 0) return;
            var off = new ulong[nb + 1];
            for (uint i = 0; i <= nb; ++i) off[i] = (ulong)Math.Min(all.Length, i * blockSize);
            var outBuf = new byte[all.Length + all.Length / 4 + 8192 * (nb + 1) + 65536];
            var outOff = new ulong[nb + 1];
            fixed (byte* pin = all, pout = outBuf)
            fixed (ulong* poff = off, pooff = outOff)
                Check(ZpaqB200Native.zpq_compress_blocks(Ctx(), pin, poff, nb, method, filename, comment, dosha1 ? 1 : 0,
                                                          pout, (ulong)outBuf.Length, pooff));
            for (ulong i = 0; i < outOff[nb]; ++i) output.put(outBuf[i]);
        }

        /// <summary>LibZPAQ.decompress (LibZPAQ.cs:65-79): every block and segment of the archive, in order.</summary>
        public static void Decompress(Reader input, Writer output)
        {
            var data = new System.IO.MemoryStream();
            int c;
            while ((c = input.get()) >= 0) data.WriteByte((byte)c);
            byte[] arc = data.ToArray();
            fixed (byte* p = arc)
            {
                long nb = ZpaqB200Native.zpq_find_blocks(p, (ulong)arc.Length, null, 0);
                if (nb < 0) LibZPAQ.error(Marshal.PtrToStringAnsi(ZpaqB200Native.zpq_last_error(IntPtr.Zero)));
                if (nb == 0) return;
                var pairs = new ulong[2 * nb];
                fixed (ulong* pp = pairs) ZpaqB200Native.zpq_find_blocks(p, (ulong)arc.Length, pp, (ulong)nb);
                // blocks are decoded from their "zPQ"; pack them back to back for the batch call
                var packed = new System.IO.MemoryStream();
                var off = new ulong[nb + 1];
                for (long i = 0; i < nb; ++i)
                {
                    off[i] = (ulong)packed.Length;
                    packed.Write(arc, (int)pairs[2 * i], (int)(pairs[2 * i + 1] - pairs[2 * i]));
                }
                off[nb] = (ulong)packed.Length;
                byte[] blocks = packed.ToArray();
                fixed (byte* pb = blocks)
                fixed (ulong* poff = off)
                {
                    long bound = ZpaqB200Native.zpq_decompressed_bound(pb, poff, (uint)nb);
                    var outBuf = new byte[Math.Max(bound, 16)];
                    var outOff = new ulong[nb + 1];
                    fixed (byte* po = outBuf)
                    fixed (ulong* pooff = outOff)
                        Check(ZpaqB200Native.zpq_decompress_blocks(Ctx(), pb, poff, (uint)nb, po, (ulong)outBuf.Length, pooff, null, null));
                    for (ulong i = 0; i < outOff[nb]; ++i) output.put(outBuf[i]);
                }
            }
        }
    }
}
