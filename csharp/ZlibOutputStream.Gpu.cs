// ZlibOutputStream.Gpu.cs -- the write side of the shim: replaces WriteCore / Finish of ZlibOutputStream.cs:125-168,213-256 when
// libzsb200 is present.  Mirrors zlibstream_b200/streams.py (the Python host the tests drive); not compiled here (no .NET).
// <copyright file="ZlibOutputStream.Gpu.cs">Licensed under the Apache License, Version 2.0.</copyright>

using System;
using System.Buffers;
using System.Collections.Generic;
using System.IO;
using SixLabors.ZlibStream.Native;

namespace SixLabors.ZlibStream
{
    /// <summary>
    /// One <c>Write</c> with <see cref="FlushMode.FullFlush"/> is one segment (ZlibOutputStream.cs:146).  A single segment gives a
    /// GPU nothing to parallelise, so writes are staged in one pooled buffer and submitted in batches; compressed bytes reach
    /// <c>BaseStream</c> in order at batch completion, on <c>Flush()</c> and on <c>Dispose()</c>.  <c>TotalOut</c> after a
    /// submitted batch and <c>SegmentEndOffsets[k]</c> are the reference's <c>TotalOut</c> after the k-th Write.
    /// </summary>
    public sealed unsafe partial class ZlibOutputStream
    {
        private IntPtr ctx;
        private byte[] stage;              // the writes of the current batch, back to back (ArrayPool)
        private int staged, stagedSegments;
        private int segSize;
        private bool shortSeen, failed, finished;
        private long submittedSegments, totalIn, totalOut;
        private uint adler = 1;
        private byte[] lastTail;           // last 32 KiB submitted: the halo of the next batch (Rle looks one byte back, SURVEY Q14)
        private int batchSegments = 4096;

        public List<long> SegmentEndOffsets { get; } = new List<long>();

        private void GpuWriteCore(ReadOnlySpan<byte> buffer)
        {
            if (buffer.IsEmpty) { return; }                                   // ZlibOutputStream.cs:127-130
            if (this.finished || this.failed) { ThrowHelper.ThrowCompressionException(true, "stream error"); }
            FlushMode mode = this.Options.FlushMode;
            if (mode == FlushMode.PartialFlush) { ThrowHelper.ThrowCompressionException(true, "FlushMode.PartialFlush is not built on the device path"); }
            if (mode != FlushMode.NoFlush)
            {
                if (this.shortSeen) { ThrowHelper.ThrowCompressionException(true, "only the last Write may be shorter than the segment size"); }
                if (this.segSize == 0)
                {
                    int n = buffer.Length;
                    this.segSize = (n >= 65536 && n % 32768 == 0) ? n : Math.Max(65536, (n + 32767) / 32768 * 32768);
                }

                if (buffer.Length > this.segSize) { ThrowHelper.ThrowCompressionException(true, "Write larger than the segment size"); }
                if (buffer.Length < this.segSize) { this.shortSeen = true; }
                this.stagedSegments++;
            }

            this.EnsureStage(this.staged + buffer.Length);
            buffer.CopyTo(this.stage.AsSpan(this.staged));                    // the one copy of the caller's bytes
            this.staged += buffer.Length;
            if (mode == FlushMode.FullFlush && this.stagedSegments >= this.batchSegments) { this.SubmitBatch(); }
        }

        private void SubmitBatch()
        {
            if (this.staged == 0 || this.failed) { return; }
            bool first = this.submittedSegments == 0;
            int cap = (int)Zsb200.zsb_deflate_bound((UIntPtr)this.staged, (UIntPtr)this.segSize);
            byte[] outBuf = ArrayPool<byte>.Shared.Rent(cap);
            ulong[] ends = ArrayPool<ulong>.Shared.Rent(this.stagedSegments);
            try
            {
                UIntPtr outLen;
                uint batchAdler;
                fixed (byte* pin = this.stage, pout = outBuf, halo = this.lastTail)
                fixed (ulong* pends = ends)
                {
                    uint flags = (first ? 0u : Zsb200.NoHeader) | Zsb200.NoTrailer;
                    int rc = Zsb200.zsb_deflate_segments(this.ctx, pin, (UIntPtr)this.staged, (UIntPtr)this.segSize,
                        (int)this.Options.CompressionLevel, (int)this.Options.CompressionStrategy, Zsb200.HashCrc32C,
                        (ulong)this.submittedSegments, first ? null : halo, flags,
                        pout, (UIntPtr)cap, out outLen, pends, out batchAdler);
                    if (rc != 0) { this.failed = true; }                      // nothing more is compressed or written
                    Zsb200.Check(this.ctx, rc, isCompressing: true);
                }

                this.BaseStream.Write(outBuf, 0, (int)outLen);                // ZlibOutputStream.cs:154
                for (int k = 0; k < this.stagedSegments; k++) { this.SegmentEndOffsets.Add(this.totalOut + (long)ends[k]); }
                this.totalOut += (long)outLen;                                // TotalOut, ZlibOutputStream.cs:73-78
                this.adler = Zsb200.zsb_adler32_combine(this.adler, batchAdler, (ulong)this.staged);
                this.totalIn += this.staged;
                this.submittedSegments += this.stagedSegments;
                this.KeepTail();
                this.staged = 0;
                this.stagedSegments = 0;
            }
            finally
            {
                ArrayPool<byte>.Shared.Return(outBuf);
                ArrayPool<ulong>.Shared.Return(ends);
            }
        }

        /// <summary>ZlibOutputStream.Finish (ZlibOutputStream.cs:213-256).</summary>
        private void GpuFinish()
        {
            if (this.finished || this.failed) { return; }
            FlushMode mode = this.Options.FlushMode;
            int level = this.Options.CompressionLevel == CompressionLevel.DefaultCompression ? 6 : (int)this.Options.CompressionLevel;
            if (mode == FlushMode.NoFlush || (mode == FlushMode.SyncFlush && this.staged > 0))
            {
                // one call for the whole stream: header, blocks, final block, Adler-32
                uint flags = mode == FlushMode.NoFlush ? Zsb200.SingleStream : Zsb200.SyncFlush;
                int seg = mode == FlushMode.NoFlush ? 0 : this.segSize;
                int cap = (int)Zsb200.zsb_deflate_bound((UIntPtr)this.staged, (UIntPtr)65536) + 64;
                byte[] outBuf = ArrayPool<byte>.Shared.Rent(cap);
                try
                {
                    fixed (byte* pin = this.stage, pout = outBuf)
                    {
                        int rc = Zsb200.zsb_deflate_segments(this.ctx, pin, (UIntPtr)this.staged, (UIntPtr)seg, level,
                            (int)this.Options.CompressionStrategy, Zsb200.HashCrc32C, 0, null, flags, pout, (UIntPtr)cap,
                            out UIntPtr outLen, null, out uint a);
                        if (rc != 0) { this.failed = true; }
                        Zsb200.Check(this.ctx, rc, isCompressing: true);
                        this.BaseStream.Write(outBuf, 0, (int)outLen);
                        this.totalOut += (long)outLen;
                        this.totalIn += this.staged;
                        this.adler = a;
                    }
                }
                finally
                {
                    ArrayPool<byte>.Shared.Return(outBuf);
                }
            }
            else
            {
                this.SubmitBatch();
                Span<byte> tail = stackalloc byte[11];
                int t = 0;
                if (this.submittedSegments == 0)
                {
                    int flevel = Math.Min(3, ((level - 1) & 0xFF) >> 1);      // Deflate.cs:466-480
                    int hdr = (0x78 << 8) | (flevel << 6);
                    hdr += 31 - (hdr % 31);
                    tail[t++] = (byte)(hdr >> 8);
                    tail[t++] = (byte)hdr;
                }

                if (level == 0) { tail[t++] = 1; tail[t++] = 0; tail[t++] = 0; tail[t++] = 0xFF; tail[t++] = 0xFF; }   // empty stored final block
                else { tail[t++] = 3; tail[t++] = 0; }                                                                // empty static final block (Q12)
                tail[t++] = (byte)(this.adler >> 24);
                tail[t++] = (byte)(this.adler >> 16);
                tail[t++] = (byte)(this.adler >> 8);
                tail[t++] = (byte)this.adler;
                this.BaseStream.Write(tail.Slice(0, t));
                this.totalOut += t;
            }

            this.finished = true;
            this.BaseStream.Flush();
        }

        private void EnsureStage(int size)
        {
            if (this.stage != null && this.stage.Length >= size) { return; }
            byte[] bigger = ArrayPool<byte>.Shared.Rent(Math.Max(size, this.stage == null ? 1 << 20 : this.stage.Length * 2));
            if (this.stage != null)
            {
                Buffer.BlockCopy(this.stage, 0, bigger, 0, this.staged);
                ArrayPool<byte>.Shared.Return(this.stage);
            }

            this.stage = bigger;
        }

        private void KeepTail()
        {
            if (this.lastTail == null) { this.lastTail = new byte[32768]; }
            if (this.staged >= 32768) { Buffer.BlockCopy(this.stage, this.staged - 32768, this.lastTail, 0, 32768); }
            else
            {
                Buffer.BlockCopy(this.lastTail, this.staged, this.lastTail, 0, 32768 - this.staged);
                Buffer.BlockCopy(this.stage, 0, this.lastTail, 32768 - this.staged, this.staged);
            }
        }
    }
}
